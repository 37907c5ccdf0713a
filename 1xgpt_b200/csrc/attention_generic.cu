// Generic CUDA-core attention (fp32 math) over the fused QKV projection output.
// Reference semantics: genie/attention.py:36-61 (BasicSelfAttention.forward):
//   q,k (optionally LayerNorm over head_dim with ONE shared affine), q *= scale,
//   softmax(q k^T  [+ causal mask value -FLT_MAX]) v.
// Used for (a) arbitrary small shapes (the SelfAttention.forward contract / test_attention.py cases),
// (b) the fp32 / tf32 parity modes, (c) cross-checking the tensor-core attention kernels.
// One CTA per (sequence, head); K and V of the sequence are staged in shared memory as fp32.
#include "kernels.cuh"
#include <cfloat>

namespace gn {
namespace {

struct GenericMap {
  int inner;                 // seq -> (b = seq / inner, s = seq % inner)
  int64_t q_outer, q_tok;    // fresh row = b*q_outer + s + i*q_tok
  int64_t c_outer, c_inner, c_tok;  // cache row = b*c_outer + s*c_inner + j*c_tok
};

template <typename T>
__device__ __forceinline__ void load_head_row(const T* p, int hd, int lane, float (&v)[4]) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int c = lane + 32 * m;
    v[m] = c < hd ? to_f32<T>(p[c]) : 0.f;
  }
}

__device__ __forceinline__ void head_layernorm(float (&v)[4], int hd, int lane, const float* g, const float* b) {
  float s = 0.f;
#pragma unroll
  for (int m = 0; m < 4; ++m) s += (lane + 32 * m < hd) ? v[m] : 0.f;
  const float mean = warp_sum(s) / (float)hd;
  float sq = 0.f;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const float dlt = (lane + 32 * m < hd) ? v[m] - mean : 0.f;
    sq += dlt * dlt;
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)hd + 1e-5f);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int c = lane + 32 * m;
    if (c < hd) v[m] = (v[m] - mean) * rstd * g[c] + b[c];
  }
}

template <typename T>
__global__ void __launch_bounds__(128)
generic_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, const T* __restrict__ kcache_in,
                         const T* __restrict__ vcache_in, T* __restrict__ kcache_out, T* __restrict__ vcache_out,
                         GenericMap mp, int n_heads, int hd, int nq, int nk_cache, int causal, float scale,
                         const float* __restrict__ qk_gamma, const float* __restrict__ qk_beta, int round_tf32) {
  extern __shared__ float sm[];
  const int nk = nk_cache + nq;
  const int d = n_heads * hd;
  const int ldk = hd + 1;
  float* sk = sm;                       // [nk][hd+1]
  float* sv = sk + (size_t)nk * ldk;    // [nk][hd]
  float* sp = sv + (size_t)nk * hd;     // [4][nk]
  float* sq = sp + 4 * (size_t)nk;      // [4][hd]
  const int seq = blockIdx.x, h = blockIdx.y;
  const int b = seq / mp.inner, s = seq % mp.inner;
  const int64_t qbase = (int64_t)b * mp.q_outer + s;
  const int64_t cbase = (int64_t)b * mp.c_outer + (int64_t)s * mp.c_inner;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int j = warp; j < nk; j += 4) {
    float kv[4], vv[4];
    if (j < nk_cache) {
      const int64_t r = cbase + (int64_t)j * mp.c_tok;
      load_head_row<T>(kcache_in + r * d + h * hd, hd, lane, kv);
      load_head_row<T>(vcache_in + r * d + h * hd, hd, lane, vv);
    } else {
      const int64_t r = qbase + (int64_t)(j - nk_cache) * mp.q_tok;
      load_head_row<T>(qkv + r * 3 * d + d + h * hd, hd, lane, kv);
      load_head_row<T>(qkv + r * 3 * d + 2 * d + h * hd, hd, lane, vv);
      if (kcache_out != nullptr) {
        const int64_t rc = cbase + (int64_t)j * mp.c_tok;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int c = lane + 32 * m;
          if (c < hd) {
            kcache_out[rc * d + h * hd + c] = from_f32<T>(kv[m]);
            vcache_out[rc * d + h * hd + c] = from_f32<T>(vv[m]);
          }
        }
      }
    }
    if (qk_gamma != nullptr) head_layernorm(kv, hd, lane, qk_gamma, qk_beta);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int c = lane + 32 * m;
      if (c < hd) {
        sk[(size_t)j * ldk + c] = kv[m];
        sv[(size_t)j * hd + c] = vv[m];
      }
    }
  }
  __syncthreads();

  float* myp = sp + (size_t)warp * nk;
  float* myq = sq + (size_t)warp * hd;
  for (int i = warp; i < nq; i += 4) {
    const int64_t r = qbase + (int64_t)i * mp.q_tok;
    float qv[4];
    load_head_row<T>(qkv + r * 3 * d + h * hd, hd, lane, qv);
    if (qk_gamma != nullptr) head_layernorm(qv, hd, lane, qk_gamma, qk_beta);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int c = lane + 32 * m;
      if (c < hd) myq[c] = qv[m] * scale;
    }
    __syncwarp();
    float mx = -FLT_MAX;
    for (int j = lane; j < nk; j += 32) {
      float acc = 0.f;
      for (int c = 0; c < hd; ++c) acc = fmaf(myq[c], sk[(size_t)j * ldk + c], acc);
      if (causal && j > i + nk_cache) acc = -FLT_MAX;
      myp[j] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < nk; j += 32) {
      const float e = expf(myp[j] - mx);
      myp[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    const int jend = causal ? min(nk, i + nk_cache + 1) : nk;
    for (int c = lane; c < hd; c += 32) {
      float acc = 0.f;
      for (int j = 0; j < jend; ++j) acc = fmaf(myp[j], sv[(size_t)j * hd + c], acc);
      float ov = acc * inv;
      if (sizeof(T) == 4 && round_tf32) ov = tf32_rn(ov);
      out[r * d + h * hd + c] = from_f32<T>(ov);
    }
    __syncwarp();
  }
}

template <typename T>
int launch_generic_t(const AttnArgs& a, int n_seq, GenericMap mp, int nq, int nk_cache, int causal, const void* kc_in,
                     const void* vc_in, void* kc_out, void* vc_out, cudaStream_t st) {
  const int nk = nq + nk_cache, hd = a.head_dim;
  GN_REQUIRE(hd <= 128, "generic attention: head_dim %d > 128", hd);
  const size_t smem = ((size_t)nk * (hd + 1) + (size_t)nk * hd + 4 * (size_t)nk + 4 * (size_t)hd) * sizeof(float);
  GN_REQUIRE(smem <= 227 * 1024, "generic attention: sequence %d x head_dim %d does not fit shared memory", nk, hd);
  auto kern = generic_attention_kernel<T>;
  static DevSmemOptIn optin;
  if (smem > 48 * 1024) GN_CUDA_CHECK(ensure_smem_optin(optin, kern, (int)smem));
  dim3 grid(n_seq, a.n_heads);
  kern<<<grid, 128, smem, st>>>(static_cast<const T*>(a.qkv), static_cast<T*>(a.out), static_cast<const T*>(kc_in),
                                static_cast<const T*>(vc_in), static_cast<T*>(kc_out), static_cast<T*>(vc_out), mp,
                                a.n_heads, hd, nq, nk_cache, causal, a.scale, a.qk_gamma, a.qk_beta, a.round_tf32);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

}  // namespace

int launch_generic_attention(const AttnArgs& a, int n_seq, int n_tok, int causal, cudaStream_t st) {
  GenericMap mp{1, n_tok, 1, 0, 0, 0};
  if (a.act_bf16 && a.fp16)
    return launch_generic_t<f16>(a, n_seq, mp, n_tok, 0, causal, nullptr, nullptr, nullptr, nullptr, st);
  return a.act_bf16 ? launch_generic_t<bf16>(a, n_seq, mp, n_tok, 0, causal, nullptr, nullptr, nullptr, nullptr, st)
                    : launch_generic_t<float>(a, n_seq, mp, n_tok, 0, causal, nullptr, nullptr, nullptr, nullptr, st);
}

int generic_temporal_attention(const AttnArgs& a, int B, int S, int T, int t0, int Tq, void* kcache, void* vcache,
                               cudaStream_t st) {
  GenericMap mp{S, (int64_t)Tq * S, S, (int64_t)S * T, T, 1};   // cache layout [B, S, T, d]
  GN_REQUIRE(t0 == 0 || (kcache && vcache), "temporal attention with t0 > 0 needs the K/V caches");
  if (a.act_bf16 && a.fp16) return launch_generic_t<f16>(a, B * S, mp, Tq, t0, 1, kcache, vcache, kcache, vcache, st);
  return a.act_bf16 ? launch_generic_t<bf16>(a, B * S, mp, Tq, t0, 1, kcache, vcache, kcache, vcache, st)
                    : launch_generic_t<float>(a, B * S, mp, Tq, t0, 1, kcache, vcache, kcache, vcache, st);
}

}  // namespace gn
