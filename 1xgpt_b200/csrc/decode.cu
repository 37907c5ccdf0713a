// MaskGIT decode-step kernels and the teacher-forced cross-entropy reduction.
// Integer work (ids, masks, ranks, scatter) is exact; softmax statistics are fp32 with warp shuffles.
#include "kernels.cuh"
#include <cfloat>

namespace gn {

// -------------------------------------------------------------------------------------
// factored softmax -> sample + confidence, one warp per token
// reference: genie/st_mask_git.py:171-190
//   probs = softmax over each 512-way vocab; id = sum_i sample_i * V^i (high vocab first);
//   conf = prod_i probs_i[sample_i]
//   temperature <= 1e-8 (uniform == nullptr): sample_i = argmax (ties: lowest index, like torch.argmax)
//   temperature  > 1e-8 (uniform != nullptr): sample_i ~ Categorical(probs_i / temperature).  Categorical
//     renormalises its `probs` argument, so the temperature cancels and the draw is from softmax(logits_i)
//     (st_mask_git.py:184-187).  The draw is made by inverse CDF from a caller-supplied uniform u in [0,1):
//     sample_i = min{ c : sum_{j<=c} e_j > u * sum_j e_j },  e_j = exp(l_j - max)  (index order, fp32).
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sample_kernel(const float* __restrict__ logits, int R, int V, int NV, const float* __restrict__ uniform,
              int32_t* __restrict__ samples, float* __restrict__ conf) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  const float* lr = logits + (int64_t)row * V * NV;
  int id = 0;
  float cf = 1.f;
  for (int i = NV - 1; i >= 0; --i) {
    const float* l = lr + (int64_t)i * V;
    float mx = -FLT_MAX;
    int arg = 0x7fffffff;
    for (int c = lane; c < V; c += 32) {
      const float v = l[c];
      if (v > mx) { mx = v; arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float sum = 0.f;
    for (int c = lane; c < V; c += 32) sum += expf(l[c] - mx);
    sum = warp_sum(sum);
    float pe = 1.f;                       // e_sample (greedy: exp(0))
    if (uniform != nullptr) {
      const float target = uniform[(int64_t)row * NV + i] * sum;
      float carry = 0.f;
      int found = -1;
      for (int c0 = 0; c0 < V && found < 0; c0 += 32) {
        const int c = c0 + lane;
        const float ev = c < V ? expf(l[c] - mx) : 0.f;
        float sc = ev;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float t = __shfl_up_sync(0xffffffffu, sc, o);
          if (lane >= o) sc += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, c < V && carry + sc > target);
        if (hit) {
          const int src = __ffs(hit) - 1;
          found = c0 + src;
          pe = __shfl_sync(0xffffffffu, ev, src);
        }
        carry += __shfl_sync(0xffffffffu, sc, 31);
      }
      if (found < 0) {                    // u * sum rounded past the last partial sum
        found = V - 1;
        pe = expf(l[V - 1] - mx);
      }
      arg = found;
    }
    id = id * V + arg;
    cf *= pe / sum;
  }
  if (lane == 0) {
    samples[row] = id;
    conf[row] = cf;
  }
}

int launch_sample(const float* logits, int R, int V, int NV, const float* uniform, int32_t* samples, float* conf,
                  cudaStream_t st) {
  const int wpb = 8;
  sample_kernel<<<ceil_div(R, wpb), wpb * 32, 0, st>>>(logits, R, V, NV, uniform, samples, conf);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// one MaskGIT re-mask step for frame out_t, one CTA per clip (st_mask_git.py:192-223)
//   prev_unmasked = unmasked;  prev = prompt[b, out_t]
//   if not last step: c = conf_or_noise; c[unmasked] = +inf; order = stable argsort(c)
//                     unmasked[order[n:]] = 1; samples[order[:n]] = mask_id
//   samples[prev_unmasked] = prev[prev_unmasked];  prompt[b, out_t] = samples
// rank_i = #{j : c_j < c_i or (c_j == c_i and j < i)}  == position of i in the stable ascending sort.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
remask_kernel(int32_t* __restrict__ prompt_frame, int64_t clip_stride, const int32_t* __restrict__ samples,
              const float* __restrict__ conf, uint8_t* __restrict__ unmasked, int32_t* __restrict__ samples_out, int S,
              int n_mask, int last_step, int mask_id) {
  extern __shared__ float sc[];
  const int b = blockIdx.x;
  int32_t* frame = prompt_frame + (int64_t)b * clip_stride;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const bool um = unmasked[(int64_t)b * S + i] != 0;
    sc[i] = (!last_step) ? (um ? INFINITY : conf[(int64_t)b * S + i]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const bool prev_um = unmasked[(int64_t)b * S + i] != 0;
    int32_t tok = samples[(int64_t)b * S + i];
    if (!last_step) {
      const float ci = sc[i];
      int rank = 0;
      for (int j = 0; j < S; ++j) {
        const float cj = sc[j];
        rank += (cj < ci || (cj == ci && j < i)) ? 1 : 0;
      }
      if (rank < n_mask) tok = mask_id;
      else unmasked[(int64_t)b * S + i] = 1;
    }
    if (prev_um) tok = frame[i];
    frame[i] = tok;
    samples_out[(int64_t)b * S + i] = tok;
  }
}

int launch_remask(int32_t* prompt_frame, int64_t clip_stride, const int32_t* samples, const float* conf_or_noise,
                  uint8_t* unmasked, int32_t* samples_out, int B, int S, int n_mask, int last_step, int mask_id,
                  cudaStream_t st) {
  GN_REQUIRE(last_step || conf_or_noise != nullptr, "remask: confidences/noise required before the last step");
  const int threads = S >= 1024 ? 1024 : ((S + 31) / 32) * 32;
  remask_kernel<<<B, threads, S * sizeof(float), st>>>(prompt_frame, clip_stride, samples, conf_or_noise, unmasked,
                                                        samples_out, S, n_mask, last_step, mask_id);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// factored CE + argmax accuracy: one warp per logits row
//   loss_row = sum_i (logsumexp(l_i) - l_i[target_i]);   ok_row = all_i (argmax l_i == target_i)
// reference: eval_utils.py:72-77 (mean taken by the caller as acc[0]/acc[1]),
//            genie/st_mask_git.py:236-250 (masked mean via `weight`)
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ce_kernel(const float* __restrict__ logits, const int32_t* __restrict__ targets, int64_t target_stride_b,
          int rows_per_b, int R, int V, int NV, const uint8_t* __restrict__ weight, double* __restrict__ acc) {
  __shared__ double s_loss[8];
  __shared__ int s_cnt[8], s_ok[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double loss = 0.0;
  int cnt = 0, okc = 0;
  for (int row = blockIdx.x * 8 + warp; row < R; row += gridDim.x * 8) {
    const int b = row / rows_per_b, r = row % rows_per_b;
    if (weight != nullptr && weight[row] == 0) continue;
    int tgt = targets[(int64_t)b * target_stride_b + r];
    const float* lr = logits + (int64_t)row * V * NV;
    float row_loss = 0.f;
    bool ok = true;
    for (int i = 0; i < NV; ++i) {
      const int ti = tgt % V;
      tgt /= V;
      const float* l = lr + (int64_t)i * V;
      float mx = -FLT_MAX;
      int arg = 0x7fffffff;
      for (int c = lane; c < V; c += 32) {
        const float v = l[c];
        if (v > mx) { mx = v; arg = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
      }
      float sum = 0.f;
      for (int c = lane; c < V; c += 32) sum += expf(l[c] - mx);
      sum = warp_sum(sum);
      row_loss += (mx + logf(sum)) - l[ti];
      ok = ok && (arg == ti);
    }
    loss += (double)row_loss;
    cnt += 1;
    okc += ok ? 1 : 0;
  }
  if (lane == 0) { s_loss[warp] = loss; s_cnt[warp] = cnt; s_ok[warp] = okc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double L = 0.0;
    int C = 0, K = 0;
    for (int w = 0; w < 8; ++w) { L += s_loss[w]; C += s_cnt[w]; K += s_ok[w]; }
    if (C) {
      atomicAdd(&acc[0], L);
      atomicAdd(&acc[1], (double)C);
      atomicAdd(&acc[2], (double)K);
    }
  }
}

int launch_ce(const float* logits, const int32_t* targets, int64_t target_stride_b, int rows_per_b, int R, int V, int NV,
              const uint8_t* weight, double* acc, cudaStream_t st) {
  int grid = ceil_div(R, 8);
  if (grid > 148 * 8) grid = 148 * 8;
  ce_kernel<<<grid, 256, 0, st>>>(logits, targets, target_stride_b, rows_per_b, R, V, NV, weight, acc);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

__global__ void __launch_bounds__(256)
count_equal_kernel(const int32_t* __restrict__ a, int64_t a_stride_b, const int32_t* __restrict__ b,
                   int64_t b_stride_b, int rows_per_b, int R, double* __restrict__ acc) {
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += gridDim.x * blockDim.x) {
    const int bb = i / rows_per_b, r = i % rows_per_b;
    c += a[(int64_t)bb * a_stride_b + r] == b[(int64_t)bb * b_stride_b + r] ? 1 : 0;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int sc[8];
  if ((threadIdx.x & 31) == 0) sc[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += sc[w];
    if (t) atomicAdd(&acc[3], (double)t);
  }
}

int launch_count_equal(const int32_t* a, int64_t a_stride_b, const int32_t* b, int64_t b_stride_b, int rows_per_b, int R,
                       double* acc, cudaStream_t st) {
  int grid = ceil_div(R, 256);
  if (grid > 592) grid = 592;
  count_equal_kernel<<<grid, 256, 0, st>>>(a, a_stride_b, b, b_stride_b, rows_per_b, R, acc);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

// -------------------------------------------------------------------------------------
// device-side replacement of the reference's host-synchronising assert (st_mask_git.py:155):
// flag |= any(prompt[:, t_from:] != mask_id)
// -------------------------------------------------------------------------------------
__global__ void check_masked_kernel(const int32_t* __restrict__ prompt, int B, int T, int S, int t_from, int mask_id,
                                    int* __restrict__ flag) {
  const int64_t per = (int64_t)(T - t_from) * S;
  const int64_t total = (int64_t)B * per;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per, r = i % per;
    bad |= prompt[(b * T + t_from) * S + r] != mask_id;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

int launch_check_masked(const int32_t* prompt, int B, int T, int S, int t_from, int mask_id, int* flag,
                        cudaStream_t st) {
  if (t_from >= T) return GN_OK;
  const int64_t total = (int64_t)B * (T - t_from) * S;
  int grid = (int)(ceil_div64(total, 256) < 592 ? ceil_div64(total, 256) : 592);
  check_masked_kernel<<<grid, 256, 0, st>>>(prompt, B, T, S, t_from, mask_id, flag);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

}  // namespace gn
