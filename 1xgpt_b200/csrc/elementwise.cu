// HBM-bound row kernels of the GENIE forward: factorized-embedding gather (+ positional add),
// LayerNorm / cast "prep" that feeds the GEMM A operand, frame gather for the readout.
// One warp per token row, 16-byte accesses, fp32 statistics via warp shuffles.
#include "kernels.cuh"
#include <cmath>
#include <type_traits>

namespace gn {

// -------------------------------------------------------------------------------------
// x[n, :] = (id == mask_id ? mask_embed : sum_i E_i[(id / V^i) % V]) + pos[t, s, :]
// reference: genie/factorization_utils.py:29-52 + genie/st_mask_git.py:261
// rows are the compact active set (b, tl in [0, Tact), s);  ids come from the full [B, T, S] window.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_kernel(const int32_t* __restrict__ ids, const float* __restrict__ E, const float* __restrict__ mask_embed,
             const float* __restrict__ pos, float* __restrict__ x, int n_rows, int d, int T, int S, int t0, int Tact,
             int V, int NV, int mask_id) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const int s = row % S;
  const int tl = (row / S) % Tact;
  const int b = row / (S * Tact);
  const int t = t0 + tl;
  const int id = ids[((int64_t)b * T + t) * S + s];
  const float4* p4 = reinterpret_cast<const float4*>(pos + ((int64_t)t * S + s) * d);
  float4* x4 = reinterpret_cast<float4*>(x + (int64_t)row * d);
  const int nvec = d >> 2;
  if (id == mask_id) {
    const float4* m4 = reinterpret_cast<const float4*>(mask_embed);
    for (int c = lane; c < nvec; c += 32) {
      const float4 m = __ldg(m4 + c), p = __ldg(p4 + c);
      x4[c] = make_float4(m.x + p.x, m.y + p.y, m.z + p.z, m.w + p.w);
    }
  } else {
    for (int c = lane; c < nvec; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int rem = id;
      for (int i = 0; i < NV; ++i) {
        const int f = rem % V;
        rem /= V;
        const float4 e = __ldg(reinterpret_cast<const float4*>(E + ((int64_t)i * V + f) * d) + c);
        acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
      }
      const float4 p = __ldg(p4 + c);
      x4[c] = make_float4(acc.x + p.x, acc.y + p.y, acc.z + p.z, acc.w + p.w);
    }
  }
}

int launch_embed(const int32_t* ids, const float* E, const float* mask_embed, const float* pos, float* x, int B,
                 int T, int S, int t0, int Tact, int d, int V, int NV, int mask_id, cudaStream_t st) {
  GN_REQUIRE(d % 4 == 0, "embed: d_model must be a multiple of 4");
  const int n_rows = B * Tact * S;
  const int wpb = 8;
  embed_kernel<<<ceil_div(n_rows, wpb), wpb * 32, 0, st>>>(ids, E, mask_embed, pos, x, n_rows, d, T, S, t0, Tact, V,
                                                            NV, mask_id);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// prep: out[r, :] = cast( (LN(x[src(r), :]) or x[src(r), :]) * scale )
//   LN: nn.LayerNorm(d, eps=1e-5), biased variance (st_transformer.py:44,67)
//   src(r): identity, or frame gather: r = (b, s) -> row (b * Tact + tsel) * S + s   (readout of frame out_t)
// -------------------------------------------------------------------------------------
template <typename OutT, int MAXV>
__global__ void __launch_bounds__(256)
prep_kernel(const float* __restrict__ x, OutT* __restrict__ out, const float* __restrict__ gamma,
            const float* __restrict__ beta, int n_rows, int d, float eps, float scale, int S, int Tact, int tsel,
            int round_tf32) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  int64_t src = row;
  if (tsel >= 0) src = ((int64_t)(row / S) * Tact + tsel) * S + (row % S);
  const float4* x4 = reinterpret_cast<const float4*>(x + src * d);
  const int nvec = d >> 2;
  float4 v[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      v[i] = x4[c];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (gamma != nullptr) {
    mean = warp_sum(sum) / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, c2 = v[i].z - mean, e = v[i].w - mean;
        sq += (a * a + b * b) + (c2 * c2 + e * e);
      }
    }
    rstd = rsqrtf(warp_sum(sq) / (float)d + eps);
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float4 o = v[i];
      if (gamma != nullptr) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
        o.x = (o.x - mean) * rstd * g.x + bt.x;
        o.y = (o.y - mean) * rstd * g.y + bt.y;
        o.z = (o.z - mean) * rstd * g.z + bt.z;
        o.w = (o.w - mean) * rstd * g.w + bt.w;
      }
      o.x *= scale; o.y *= scale; o.z *= scale; o.w *= scale;
      if (sizeof(OutT) == 4) {
        if (round_tf32) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
        reinterpret_cast<float4*>(out + (int64_t)row * d)[c] = o;
      } else {
        typedef typename std::conditional<sizeof(OutT) == 2, OutT, bf16>::type P16;
        uint2 p;
        p.x = pack_h2<P16>(o.x, o.y);
        p.y = pack_h2<P16>(o.z, o.w);
        reinterpret_cast<uint2*>(out + (int64_t)row * d)[c] = p;
      }
    }
  }
}

template <typename OutT>
static int launch_prep_t(const float* x, OutT* out, const float* gamma, const float* beta, int n_rows, int d,
                         float scale, int S, int Tact, int tsel, cudaStream_t st, int round_tf32) {
  // 8 warps (= rows) per block.  (Capping the resident blocks so that the last wave is full - 7 blocks per SM give
  // 3.95 / 1.98 waves instead of 3.46 / 1.73 at the 32768 / 16384 rows of a decode step - was measured SLOWER: LayerNorm
  // passes 22.7 -> 24.7 ms per step; the pass wants every resident warp it can get.)
  const int wpb = 8;
  const size_t pad = 0;
  const int grid = ceil_div(n_rows, wpb);
  const int nvec = d / 4;
  if (nvec <= 32 * 2)
    GN_CUDA_CHECK(launch_kernel(PC_PREP, prep_kernel<OutT, 2>, dim3(grid), dim3(wpb * 32), pad, st, x, out, gamma, beta, n_rows, d, 1e-5f, scale, S, Tact, tsel, round_tf32));
  else if (nvec <= 32 * 4)
    GN_CUDA_CHECK(launch_kernel(PC_PREP, prep_kernel<OutT, 4>, dim3(grid), dim3(wpb * 32), pad, st, x, out, gamma, beta, n_rows, d, 1e-5f, scale, S, Tact, tsel, round_tf32));
  else if (nvec <= 32 * 8)
    GN_CUDA_CHECK(launch_kernel(PC_PREP, prep_kernel<OutT, 8>, dim3(grid), dim3(wpb * 32), pad, st, x, out, gamma, beta, n_rows, d, 1e-5f, scale, S, Tact, tsel, round_tf32));
  else if (nvec <= 32 * 16)
    GN_CUDA_CHECK(launch_kernel(PC_PREP, prep_kernel<OutT, 16>, dim3(grid), dim3(wpb * 32), pad, st, x, out, gamma, beta, n_rows, d, 1e-5f, scale, S, Tact, tsel, round_tf32));
  else {
    set_error("prep: d_model %d > 2048 not supported", d);
    return GN_ERR_UNSUPPORTED;
  }
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

int launch_prep(const float* x, void* out, int out_bf16, const float* gamma, const float* beta, int n_rows, int d,
                float scale, int S, int Tact, int tsel, cudaStream_t st, int round_tf32) {
  GN_REQUIRE(d % 4 == 0, "prep: d_model must be a multiple of 4");
  GN_REQUIRE((gamma == nullptr) == (beta == nullptr), "prep: gamma/beta must both be set or both be null");
  if (out_bf16 == 2) return launch_prep_t<f16>(x, static_cast<f16*>(out), gamma, beta, n_rows, d, scale, S, Tact, tsel, st, 0);
  if (out_bf16) return launch_prep_t<bf16>(x, static_cast<bf16*>(out), gamma, beta, n_rows, d, scale, S, Tact, tsel, st, 0);
  return launch_prep_t<float>(x, static_cast<float*>(out), gamma, beta, n_rows, d, scale, S, Tact, tsel, st, round_tf32);
}

// bf16 cast + row statistics (feeds the folded-LayerNorm GEMM epilogue of the first layer)
__global__ void __launch_bounds__(256)
prep_stats_kernel(const float* __restrict__ x, bf16* __restrict__ out, float* __restrict__ stats, int n_rows, int d) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const float4* x4 = reinterpret_cast<const float4*>(x + (int64_t)row * d);
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < d / 4; c += 32) {
    const float4 v = x4[c];
    s1 += (v.x + v.y) + (v.z + v.w);
    s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    uint2 p;
    p.x = pack_bf16x2(v.x, v.y);
    p.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(out + (int64_t)row * d)[c] = p;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) reinterpret_cast<float2*>(stats)[row] = make_float2(s1, s2);
}
int launch_prep_stats(const float* x, bf16* out, float* stats, int n_rows, int d, cudaStream_t st) {
  GN_REQUIRE(d % 4 == 0, "prep: d_model must be a multiple of 4");
  GN_CUDA_CHECK(launch_kernel(PC_PREP, prep_stats_kernel, dim3(ceil_div(n_rows, 8)), dim3(256), 0, st, x, out, stats,
                              n_rows, d));
  ++g_launch_count;
  return GN_OK;
}

// LayerNorm folding, one warp per output feature n
__global__ void __launch_bounds__(256)
fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ bias, bf16* __restrict__ Wf, float* __restrict__ colsum,
               float* __restrict__ bias_f, int N, int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  float cs = 0.f, bs = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[(int64_t)n * K + k];
    const bf16 wf = __float2bfloat16_rn(w * gamma[k]);
    Wf[(int64_t)n * K + k] = wf;
    cs += __bfloat162float(wf);          // the sum the tensor core will see (mean subtraction stays exact)
    bs = fmaf(beta[k], w, bs);
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) {
    colsum[n] = cs;
    bias_f[n] = bs + (bias ? bias[n] : 0.f);
  }
}
int launch_fold_ln(const float* W, const float* gamma, const float* beta, const float* bias, bf16* Wf, float* colsum,
                   float* bias_f, int N, int K, cudaStream_t st) {
  fold_ln_kernel<<<ceil_div(N, 8), 256, 0, st>>>(W, gamma, beta, bias, Wf, colsum, bias_f, N, K);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// fp32 -> bf16 weight conversion (weights are repacked once at load time)
// -------------------------------------------------------------------------------------
template <typename H>
__global__ void cast_h16_kernel(const float* __restrict__ in, H* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = from_f32<H>(in[i]);
}
__global__ void round_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = tf32_rn(in[i]);
}
int launch_round_tf32(const float* in, float* out, int64_t n, cudaStream_t st) {
  if (n == 0) return GN_OK;
  const int grid = (int)(ceil_div64(n, 256) < 4096 ? ceil_div64(n, 256) : 4096);
  round_tf32_kernel<<<grid, 256, 0, st>>>(in, out, n);
  GN_CUDA_CHECK(cudaGetLastError());
  return GN_OK;
}
int launch_cast_h16(const float* in, void* out, int fp16, int64_t n, cudaStream_t st) {
  if (n == 0) return GN_OK;
  const int grid = (int)(ceil_div64(n, 256) < 4096 ? ceil_div64(n, 256) : 4096);
  if (fp16) cast_h16_kernel<f16><<<grid, 256, 0, st>>>(in, static_cast<f16*>(out), n);
  else cast_h16_kernel<bf16><<<grid, 256, 0, st>>>(in, static_cast<bf16*>(out), n);
  GN_CUDA_CHECK(cudaGetLastError());
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// logits rows [R, C] (R = (b, tl, s) compact) -> reference layout [B, C, Tout, S] at frame slot `tslot`
// (st_mask_git.py:264  "B T (H W) C -> B C T H W").  32x32 smem transpose tiles.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
logits_transpose_kernel(const float* __restrict__ rows, float* __restrict__ out, int C, int S, int Tl, int Tout,
                        int tslot0) {
  __shared__ float tile[32][33];
  // grid: x over S/32, y over C/32, z over (b * Tl + tl)
  const int bt = blockIdx.z;
  const int b = bt / Tl, tl = bt % Tl;
  const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int s = s0 + i, c = c0 + tx;
    tile[i][tx] = (s < S && c < C) ? rows[((int64_t)bt * S + s) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, s = s0 + tx;
    if (s < S && c < C) out[(((int64_t)b * C + c) * Tout + (tslot0 + tl)) * S + s] = tile[tx][i];
  }
}
int launch_logits_transpose(const float* rows, float* out, int B, int Tl, int S, int C, int Tout, int tslot0,
                            cudaStream_t st) {
  dim3 grid(ceil_div(S, 32), ceil_div(C, 32), B * Tl);
  logits_transpose_kernel<<<grid, 256, 0, st>>>(rows, out, C, S, Tl, Tout, tslot0);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

}  // namespace gn
