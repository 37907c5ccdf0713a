// bf16 attention kernels for the two production shapes of the ST block (st_transformer.py:70-83):
//   spatial : non-causal attention over the S = 256 (or 128) tokens of one frame, head_dim 64 / 32
//   temporal: causal attention over the <= 16 frames of one spatial position, with an optional K/V cache
// Both read the fused QKV projection output [rows, 3d] (column order (3, h, hd), attention.py:38)
// without any physical (B T) S C <-> (B S) T C transpose: the sequence structure is expressed through
// TMA box coordinates (spatial) or row strides (temporal).
// Tiles are staged in 128B/64B-swizzled shared memory (TMA for the spatial kernel), QK^T and PV run on
// the warp-level tensor-core path (mma.sync m16n8k16 bf16 -> f32), softmax statistics are fp32 with
// quad shuffles.  Optional qk-LayerNorm (attention.py:42-47) is applied to the staged Q/K rows in place.
#include "kernels.cuh"
#include "tensormap.cuh"

namespace gn {
namespace {

template <typename H>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (sizeof(H) == 2 && H16<H>::UMMA_FMT == 1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// byte offset of 16-byte chunk `c` of row `r` in a tile whose rows are HD bf16 wide, swizzled the way TMA
// SWIZZLE_128B (HD = 64) / SWIZZLE_64B (HD = 32) lays it out (tile base aligned to 1024 B).
template <int HD>
__device__ __forceinline__ uint32_t swz(int r, int c) {
  if (HD == 64) return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
  return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
}

// In-place LayerNorm(head_dim) of `nrows` staged rows (shared affine, eps 1e-5), CH = HD/8 lanes per row.
template <int HD, typename H>
__device__ __forceinline__ void qk_layernorm_rows(uint8_t* tile, int nrows, int tid, int nthreads,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta) {
  constexpr int CH = HD / 8;
  const int sub = tid % CH;
  // uniform trip count: the shuffles below are executed by every lane of the warp
  for (int rb = 0; rb < nrows; rb += nthreads / CH) {
    const int r = rb + tid / CH;
    const bool act = r < nrows;
    uint4* p = reinterpret_cast<uint4*>(tile + swz<HD>(act ? r : 0, sub));
    uint4 raw = act ? *p : make_uint4(0, 0, 0, 0);
    const uint32_t* h2 = reinterpret_cast<const uint32_t*>(&raw);
    float v[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_h2<H>(h2[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
    for (int o = CH / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)HD;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float dl = v[i] - mean; sq += dl * dl; }
#pragma unroll
    for (int o = CH / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)HD + 1e-5f);
    uint4 outv;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&outv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c0 = sub * 8 + 2 * i;
      ow[i] = pack_h2<H>((v[2 * i] - mean) * rstd * gamma[c0] + beta[c0],
                          (v[2 * i + 1] - mean) * rstd * gamma[c0 + 1] + beta[c0 + 1]);
    }
    if (act) *p = outv;
  }
}

// =====================================================================================
// spatial attention: grid (S/128, heads, frames), 8 warps x 16 query rows
// =====================================================================================
constexpr int SP_THREADS = 256;
constexpr int SP_QROWS = 128;

template <int HD, typename H>
__global__ void __launch_bounds__(SP_THREADS, 2)
spatial_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    H* __restrict__ out, int S, int d, float scale_log2e, const float* __restrict__ gamma,
                    const float* __restrict__ beta) {
  constexpr int ROWB = HD * 2;  // bytes per staged row
  constexpr int CH = HD / 8;    // 16-byte chunks per row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + SP_QROWS * ROWB;
  uint8_t* sV = sK + S * ROWB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + S * ROWB);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, f = blockIdx.z;
  const int frame_row0 = f * S;
  const int q_row0 = frame_row0 + qt * SP_QROWS;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], (SP_QROWS + S) * ROWB);
    tma_load_2d(sQ, &tmQ, &bars[0], h * HD, q_row0);
    tma_load_2d(sK, &tmKV, &bars[0], d + h * HD, frame_row0);
    mbar_arrive_expect_tx(&bars[1], S * ROWB);
    tma_load_2d(sV, &tmKV, &bars[1], 2 * d + h * HD, frame_row0);
  }
  mbar_wait(&bars[0], 0);
  if (gamma != nullptr) {
    qk_layernorm_rows<HD, H>(sQ, SP_QROWS, tid, SP_THREADS, gamma, beta);
    qk_layernorm_rows<HD, H>(sK, S, tid, SP_THREADS, gamma, beta);
    __syncthreads();
  }

  const uint32_t sQa = smem_u32(sQ), sKa = smem_u32(sK), sVa = smem_u32(sV);
  const int mi = lane >> 3, l8 = lane & 7;
  const int g = lane >> 2, t4 = lane & 3;

  uint32_t qf[HD / 16][4];
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk)
    ldsm_x4(qf[kk], sQa + swz<HD>(warp * 16 + l8 + (mi & 1) * 8, 2 * kk + (mi >> 1)));

  float o[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const int nkb = S / 64;
  for (int kb = 0; kb < nkb; ++kb) {
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b[4];
        ldsm_x4(b, sKa + swz<HD>(kb * 64 + (2 * jp + (mi >> 1)) * 8 + l8, 2 * kk + (mi & 1)));
        mma_16816<H>(s[2 * jp], qf[kk], b[0], b[1]);
        mma_16816<H>(s[2 * jp + 1], qf[kk], b[2], b[3]);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float a0 = exp2f((m0 - mn0) * scale_log2e), a1 = exp2f((m1 - mn1) * scale_log2e);
    m0 = mn0; m1 = mn1;
    const float off0 = mn0 * scale_log2e, off1 = mn1 * scale_log2e;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f(fmaf(s[j][0], scale_log2e, -off0));
      s[j][1] = exp2f(fmaf(s[j][1], scale_log2e, -off0));
      s[j][2] = exp2f(fmaf(s[j][2], scale_log2e, -off1));
      s[j][3] = exp2f(fmaf(s[j][3], scale_log2e, -off1));
      rs0 += s[j][0] + s[j][1];
      rs1 += s[j][2] + s[j][3];
    }
    l0 = l0 * a0 + rs0;
    l1 = l1 * a1 + rs1;
#pragma unroll
    for (int j = 0; j < HD / 8; ++j) { o[j][0] *= a0; o[j][1] *= a0; o[j][2] *= a1; o[j][3] *= a1; }
    if (kb == 0) mbar_wait(&bars[1], 0);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      uint32_t a[4];
      a[0] = pack_h2<H>(s[2 * k2][0], s[2 * k2][1]);
      a[1] = pack_h2<H>(s[2 * k2][2], s[2 * k2][3]);
      a[2] = pack_h2<H>(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
      a[3] = pack_h2<H>(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
      for (int jp = 0; jp < HD / 16; ++jp) {
        uint32_t b[4];
        ldsm_x4_t(b, sVa + swz<HD>(kb * 64 + k2 * 16 + (mi & 1) * 8 + l8, 2 * jp + (mi >> 1)));
        mma_16816<H>(o[2 * jp], a, b[0], b[1]);
        mma_16816<H>(o[2 * jp + 1], a, b[2], b[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;

  // stage the warp's 16 x HD output in its own (now dead) Q rows, then 16-byte coalesced stores
  __syncwarp();
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) {
    *reinterpret_cast<uint32_t*>(sQ + swz<HD>(warp * 16 + g, j) + t4 * 4) = pack_h2<H>(o[j][0] * i0, o[j][1] * i0);
    *reinterpret_cast<uint32_t*>(sQ + swz<HD>(warp * 16 + g + 8, j) + t4 * 4) = pack_h2<H>(o[j][2] * i1, o[j][3] * i1);
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 16 * CH / 32; ++it) {
    const int idx = it * 32 + lane;
    const int r = idx / CH, c = idx % CH;
    const uint4 v = *reinterpret_cast<const uint4*>(sQ + swz<HD>(warp * 16 + r, c));
    *reinterpret_cast<uint4*>(out + (int64_t)(q_row0 + warp * 16 + r) * d + h * HD + c * 8) = v;
  }
}

template <int HD, typename H>
int launch_spatial_t(const AttnArgs& a, int n_frames, int S, cudaStream_t st) {
  const int d = a.n_heads * a.head_dim;
  const int64_t rows = (int64_t)n_frames * S;
  CUtensorMap tmQ, tmKV;
  const CUtensorMapSwizzle sw = HD == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  GN_PROPAGATE(make_tensor_map_2d(&tmQ, a.qkv, H16<H>::TMAP, 2, 3 * d, rows, 3 * d, HD, SP_QROWS, sw));
  GN_PROPAGATE(make_tensor_map_2d(&tmKV, a.qkv, H16<H>::TMAP, 2, 3 * d, rows, 3 * d, HD, S, sw));
  const int smem = (SP_QROWS + 2 * S) * HD * 2 + 64 + 1024;
  auto kern = spatial_attn_kernel<HD, H>;
  static DevSmemOptIn optin;
  GN_CUDA_CHECK(ensure_smem_optin(optin, kern, smem));
  dim3 grid(S / SP_QROWS, a.n_heads, n_frames);
  GN_CUDA_CHECK(launch_kernel(PC_SPATIAL, kern, grid, dim3(SP_THREADS), (size_t)smem, st, tmQ, tmKV, static_cast<H*>(a.out), S, d,
                              a.scale * 1.4426950408889634f, a.qk_gamma, a.qk_beta));
  ++g_launch_count;
  return GN_OK;
}

// =====================================================================================
// temporal attention: persistent CTAs, one warp per head, looping over (clip, spatial position) with a 2-stage
// cp.async ring: the K/V rows (cache + fresh) of position i+1 stream into shared memory while position i is
// computed, so the kernel is a continuous HBM stream (it is bandwidth-bound: (t0+Tq)*2*d*2 B of K/V per token).
// =====================================================================================
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 16 : 0;   // src-size 0 -> the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int HD, typename H>
__global__ void __launch_bounds__(512)
temporal_attn_kernel(const H* __restrict__ qkv, H* __restrict__ out, H* __restrict__ kcache,
                     H* __restrict__ vcache, int n_pos, int S, int T, int t0, int Tq, int d, float scale_log2e,
                     const float* __restrict__ gamma, const float* __restrict__ beta) {
  constexpr int ROWB = HD * 2;
  constexpr int CH = HD / 8;
  constexpr int TILEB = 16 * ROWB;
  constexpr int NIT = 16 * CH / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = warp;
  uint8_t* wbase = smem + warp * 2 * 3 * TILEB;     // [stage][q|k|v]
  const int Tk = t0 + Tq;
  const int mi = lane >> 3, l8 = lane & 7;
  const int g = lane >> 2, t4 = lane & 3;
  pdl_wait();
  pdl_trigger();

  auto prefetch = [&](int pos, int stage) {
    const int b = pos / S, sp = pos % S;
    const int64_t fresh0 = ((int64_t)b * Tq) * S + sp;     // fresh row of local frame tl: fresh0 + tl*S
    const int64_t cache0 = ((int64_t)b * S + sp) * T;      // cache row of frame j (layout [B,S,T,d])
    uint8_t* sQ = wbase + stage * 3 * TILEB;
    uint8_t* sK = sQ + TILEB;
    uint8_t* sV = sK + TILEB;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = it * 32 + lane;
      const int r = idx / CH, c = idx % CH;
      const H* qsrc = qkv + (fresh0 + (int64_t)(r < Tq ? r : 0) * S) * 3 * d + h * HD + c * 8;
      cp_async16(sQ + swz<HD>(r, c), qsrc, r < Tq);
      const H *ksrc, *vsrc;
      if (r < t0) {
        const int64_t cr = (cache0 + r) * d + h * HD + c * 8;
        ksrc = kcache + cr;
        vsrc = vcache + cr;
      } else {
        const int64_t fr = (fresh0 + (int64_t)(r < Tk ? r - t0 : 0) * S) * 3 * d + h * HD + c * 8;
        ksrc = qkv + fr + d;
        vsrc = qkv + fr + 2 * d;
      }
      cp_async16(sK + swz<HD>(r, c), ksrc, r < Tk);
      cp_async16(sV + swz<HD>(r, c), vsrc, r < Tk);
    }
    cp_async_commit();
  };

  int pos = blockIdx.x;
  int stage = 0;
  if (pos < n_pos) prefetch(pos, 0);
  for (; pos < n_pos; pos += gridDim.x, stage ^= 1) {
    const int nxt = pos + gridDim.x;
    if (nxt < n_pos) {
      prefetch(nxt, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const int b = pos / S, sp = pos % S;
    const int64_t fresh0 = ((int64_t)b * Tq) * S + sp;
    const int64_t cache0 = ((int64_t)b * S + sp) * T;
    uint8_t* sQ = wbase + stage * 3 * TILEB;
    uint8_t* sK = sQ + TILEB;
    uint8_t* sV = sK + TILEB;
    // append the fresh frames (raw projections) to the K/V cache
    if (kcache != nullptr) {
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int idx = it * 32 + lane;
        const int r = idx / CH, c = idx % CH;
        if (r >= t0 && r < Tk) {
          const int64_t cr = (cache0 + r) * d + h * HD + c * 8;
          *reinterpret_cast<uint4*>(kcache + cr) = *reinterpret_cast<const uint4*>(sK + swz<HD>(r, c));
          *reinterpret_cast<uint4*>(vcache + cr) = *reinterpret_cast<const uint4*>(sV + swz<HD>(r, c));
        }
      }
    }
    if (gamma != nullptr) {
      qk_layernorm_rows<HD, H>(sQ, Tq, lane, 32, gamma, beta);
      qk_layernorm_rows<HD, H>(sK, Tk, lane, 32, gamma, beta);
      __syncwarp();
    }
    const uint32_t sQa = smem_u32(sQ), sKa = smem_u32(sK), sVa = smem_u32(sV);

    float s[2][4];
    s[0][0] = s[0][1] = s[0][2] = s[0][3] = 0.f;
    s[1][0] = s[1][1] = s[1][2] = s[1][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t qf[4], kf[4];
      ldsm_x4(qf, sQa + swz<HD>(l8 + (mi & 1) * 8, 2 * kk + (mi >> 1)));
      ldsm_x4(kf, sKa + swz<HD>((mi >> 1) * 8 + l8, 2 * kk + (mi & 1)));
      mma_16816<H>(s[0], qf, kf[0], kf[1]);
      mma_16816<H>(s[1], qf, kf[2], kf[3]);
    }
    // causal mask: query row i (frame t0 + i) sees keys j <= t0 + i   (attention.py:51-55)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = j * 8 + 2 * t4 + e;
        if (key > t0 + g) s[j][e] = -INFINITY;
        if (key > t0 + g + 8) s[j][2 + e] = -INFINITY;
        mx0 = fmaxf(mx0, s[j][e]);
        mx1 = fmaxf(mx1, s[j][2 + e]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[j][e] = exp2f((s[j][e] - mx0) * scale_log2e);
        s[j][2 + e] = exp2f((s[j][2 + e] - mx1) * scale_log2e);
        l0 += s[j][e];
        l1 += s[j][2 + e];
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    uint32_t pa[4];
    pa[0] = pack_h2<H>(s[0][0], s[0][1]);
    pa[1] = pack_h2<H>(s[0][2], s[0][3]);
    pa[2] = pack_h2<H>(s[1][0], s[1][1]);
    pa[3] = pack_h2<H>(s[1][2], s[1][3]);
    float o[HD / 8][4];
#pragma unroll
    for (int jp = 0; jp < HD / 16; ++jp) {
      uint32_t vf[4];
      ldsm_x4_t(vf, sVa + swz<HD>((mi & 1) * 8 + l8, 2 * jp + (mi >> 1)));
      o[2 * jp][0] = o[2 * jp][1] = o[2 * jp][2] = o[2 * jp][3] = 0.f;
      o[2 * jp + 1][0] = o[2 * jp + 1][1] = o[2 * jp + 1][2] = o[2 * jp + 1][3] = 0.f;
      mma_16816<H>(o[2 * jp], pa, vf[0], vf[1]);
      mma_16816<H>(o[2 * jp + 1], pa, vf[2], vf[3]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < HD / 8; ++j) {
      *reinterpret_cast<uint32_t*>(sQ + swz<HD>(g, j) + t4 * 4) = pack_h2<H>(o[j][0] * i0, o[j][1] * i0);
      *reinterpret_cast<uint32_t*>(sQ + swz<HD>(g + 8, j) + t4 * 4) = pack_h2<H>(o[j][2] * i1, o[j][3] * i1);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = it * 32 + lane;
      const int r = idx / CH, c = idx % CH;
      if (r < Tq) {
        const uint4 v = *reinterpret_cast<const uint4*>(sQ + swz<HD>(r, c));
        *reinterpret_cast<uint4*>(out + (fresh0 + (int64_t)r * S) * d + h * HD + c * 8) = v;
      }
    }
    __syncwarp();   // this stage is refilled two iterations from now
  }
}

template <int HD, typename H>
int launch_temporal_t(const AttnArgs& a, int B, int S, int T, int t0, int Tq, void* kcache, void* vcache,
                      cudaStream_t st) {
  const int d = a.n_heads * a.head_dim;
  const int smem = a.n_heads * 2 * 3 * 16 * HD * 2 + 1024;
  auto kern = temporal_attn_kernel<HD, H>;
  static DevSmemOptIn optin;
  if (smem > 48 * 1024) GN_CUDA_CHECK(ensure_smem_optin(optin, kern, smem));
  const int sms = device_sm_count();
  const int per_sm = std::max(1, (227 * 1024) / smem);
  const int n_pos = B * S;
  const int grid = std::min(n_pos, sms * per_sm);
  GN_CUDA_CHECK(launch_kernel(PC_TEMPORAL, kern, dim3(grid), dim3(a.n_heads * 32), (size_t)smem, st,
                              static_cast<const H*>(a.qkv), static_cast<H*>(a.out), static_cast<H*>(kcache),
                              static_cast<H*>(vcache), n_pos, S, T, t0, Tq, d, a.scale * 1.4426950408889634f,
                              a.qk_gamma, a.qk_beta));
  ++g_launch_count;
  return GN_OK;
}


// =====================================================================================
// temporal attention v2 (head_dim 64, no qk-LayerNorm): TMA-fed, no per-thread global access.
// K and V of ALL frames <= the query frame come from a head-major cache [clip*S + s][head][T][64] that the
// temporal QKV GEMM epilogue writes directly (gemm.cu, kv_* arguments), so one (clip, position) needs exactly
// three bulk-tensor loads (Q box [Tq][H][64] from the projection output, K and V boxes [H][Tk][64] from the
// caches) and one bulk-tensor store of the attention output.  Persistent CTAs: warp h < H computes head h
// (mma.sync m16n8k16, the 16 x 16 causal problem of one position), warp H is the TMA producer running an
// n_stages-deep mbarrier ring ahead of the consumers.  Key r of a position always sits at tile row r, whatever
// (t0, Tq) a call uses, so the cached decode stays bit-identical to the dense 16-frame forward.
// =====================================================================================
// byte offset of 16-byte chunk c of `line` (one (frame, head) row of HD elements) in a TMA-swizzled region whose base
// is aligned to 1024 B: SWIZZLE_128B for 128-byte lines (HD = 64), SWIZZLE_64B for 64-byte lines (HD = 32)
template <int HD>
__device__ __forceinline__ uint32_t line_off(int line, int c) {
  if (HD == 64) return (uint32_t)(line * 128 + ((c ^ (line & 7)) << 4));
  return (uint32_t)(line * 64 + ((c ^ ((line >> 1) & 3)) << 4));
}

template <typename E, int HD>
__global__ void __launch_bounds__(576, 1)
temporal_attn_v2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, int n_pos,
                        int S, int H, int t0, int Tq, int n_stages, int rq_bytes, int rk_bytes, float scale_log2e,
                        int hint) {
  static_assert(HD == 64 || HD == 32, "head_dim 64 or 32");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = rq_bytes + 2 * rk_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + n_stages * stage_bytes);
  uint64_t* empty_bar = full_bar + n_stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tk = t0 + Tq;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], H);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  const int n_my = ((int)blockIdx.x < n_pos) ? (n_pos - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == H) {
    // ------------------------------------------------------------------ TMA producer (one thread)
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(Tq + 2 * Tk) * H * (HD * 2);
      const uint64_t pol = l2_policy_evict_first();
      auto issue = [&](int j, int st) {
        const int pos = blockIdx.x + j * gridDim.x;
        const int b = pos / S, sp = pos - b * S;
        uint8_t* base = smem + st * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[st], tx);
        tma_load_4d(base, &tmQ, &full_bar[st], 0, 0, sp, b * Tq);
        if (hint) {   // the K/V cache is a pure stream (hundreds of MB per launch): do not let it flush L2
          tma_load_3d_hint(base + rq_bytes, &tmK, &full_bar[st], 0, 0, pos * H, pol);
          tma_load_3d_hint(base + rq_bytes + rk_bytes, &tmV, &full_bar[st], 0, 0, pos * H, pol);
        } else {
          tma_load_3d(base + rq_bytes, &tmK, &full_bar[st], 0, 0, pos * H);
          tma_load_3d(base + rq_bytes + rk_bytes, &tmV, &full_bar[st], 0, 0, pos * H);
        }
      };
      for (int j = 0; j < n_stages && j < n_my; ++j) issue(j, j);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_my; ++j) {
        mbar_wait(&empty_bar[st], ph);            // every head's output rows are staged in the Q region
        const int pos = blockIdx.x + j * gridDim.x;
        const int b = pos / S, sp = pos - b * S;
        tma_store_4d(&tmO, smem + st * stage_bytes, 0, 0, sp, b * Tq);
        tma_store_commit();
        if (j + n_stages < n_my) {
          tma_store_wait_read<0>();               // the store has read the stage: refill it
          issue(j + n_stages, st);
        }
        if (++st == n_stages) { st = 0; ph ^= 1; }
      }
      tma_store_wait_all<0>();
    }
  } else {
    // ------------------------------------------------------------------ one head per warp
    const int h = warp;
    const int mi = lane >> 3, l8 = lane & 7;
    const int g = lane >> 2, t4 = lane & 3;
    const int qrow = min(l8 + (mi & 1) * 8, Tq - 1);          // rows past the real ones re-read the last real row
    const int krow = min((mi >> 1) * 8 + l8, Tk - 1);
    const int vrow = min((mi & 1) * 8 + l8, Tk - 1);
    uint32_t qo[HD / 16], ko[HD / 16], vo[HD / 16];
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      qo[kk] = line_off<HD>(qrow * H + h, 2 * kk + (mi >> 1));
      ko[kk] = line_off<HD>(h * Tk + krow, 2 * kk + (mi & 1));
      vo[kk] = line_off<HD>(h * Tk + vrow, 2 * kk + (mi >> 1));
    }
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_my; ++j) {
      uint8_t* base = smem + st * stage_bytes;
      const uint32_t sQa = smem_u32(base), sKa = sQa + rq_bytes, sVa = sKa + rk_bytes;
      mbar_wait(&full_bar[st], ph);
      float s[2][4];
      s[0][0] = s[0][1] = s[0][2] = s[0][3] = 0.f;
      s[1][0] = s[1][1] = s[1][2] = s[1][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        uint32_t qf[4], kf[4];
        ldsm_x4(qf, sQa + qo[kk]);
        ldsm_x4(kf, sKa + ko[kk]);
        mma_16816<E>(s[0], qf, kf[0], kf[1]);
        mma_16816<E>(s[1], qf, kf[2], kf[3]);
      }
      // causal mask: query row i (frame t0 + i) sees keys j <= t0 + i   (attention.py:51-55)
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = jj * 8 + 2 * t4 + e;
          if (key > t0 + g) s[jj][e] = -INFINITY;
          if (key > t0 + g + 8) s[jj][2 + e] = -INFINITY;
          mx0 = fmaxf(mx0, s[jj][e]);
          mx1 = fmaxf(mx1, s[jj][2 + e]);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          s[jj][e] = exp2f((s[jj][e] - mx0) * scale_log2e);
          s[jj][2 + e] = exp2f((s[jj][2 + e] - mx1) * scale_log2e);
          l0 += s[jj][e];
          l1 += s[jj][2 + e];
        }
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float i0 = 1.f / l0, i1 = 1.f / l1;
      uint32_t pa[4];
      pa[0] = pack_h2<E>(s[0][0], s[0][1]);
      pa[1] = pack_h2<E>(s[0][2], s[0][3]);
      pa[2] = pack_h2<E>(s[1][0], s[1][1]);
      pa[3] = pack_h2<E>(s[1][2], s[1][3]);
      float o[HD / 8][4];
#pragma unroll
      for (int jp = 0; jp < HD / 16; ++jp) {
        uint32_t vf[4];
        ldsm_x4_t(vf, sVa + vo[jp]);
        o[2 * jp][0] = o[2 * jp][1] = o[2 * jp][2] = o[2 * jp][3] = 0.f;
        o[2 * jp + 1][0] = o[2 * jp + 1][1] = o[2 * jp + 1][2] = o[2 * jp + 1][3] = 0.f;
        mma_16816<E>(o[2 * jp], pa, vf[0], vf[1]);
        mma_16816<E>(o[2 * jp + 1], pa, vf[2], vf[3]);
      }
      __syncwarp();        // every lane's Q fragments are in registers: the Q lines of this head become output staging
      if (g < Tq) {
#pragma unroll
        for (int c = 0; c < HD / 8; ++c)
          *reinterpret_cast<uint32_t*>(base + line_off<HD>(g * H + h, c) + t4 * 4) = pack_h2<E>(o[c][0] * i0, o[c][1] * i0);
      }
      if (g + 8 < Tq) {
#pragma unroll
        for (int c = 0; c < HD / 8; ++c)
          *reinterpret_cast<uint32_t*>(base + line_off<HD>((g + 8) * H + h, c) + t4 * 4) =
              pack_h2<E>(o[c][2] * i1, o[c][3] * i1);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
      if (++st == n_stages) { st = 0; ph ^= 1; }
    }
  }
}

template <typename E, int HD>
int launch_temporal_v2(const AttnArgs& a, int nb, int S, int T, int t0, int Tq, const void* kc, const void* vc,
                       cudaStream_t st) {
  const int H = a.n_heads, d = H * HD, Tk = t0 + Tq;
  const CUtensorMapSwizzle tsw = HD == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const int n_pos = nb * S;
  CUtensorMap tmQ, tmK, tmV, tmO;
  {
    const int64_t dims[4] = {HD, H, S, (int64_t)nb * Tq};
    const int64_t sq[3] = {HD * 2, (int64_t)3 * d * 2, (int64_t)S * 3 * d * 2};
    const int64_t so[3] = {HD * 2, (int64_t)d * 2, (int64_t)S * d * 2};
    const int box[4] = {HD, H, 1, Tq};
    GN_PROPAGATE(make_tensor_map_nd(&tmQ, a.qkv, H16<E>::TMAP, 4, dims, sq, box,
                                    tsw));
    GN_PROPAGATE(make_tensor_map_nd(&tmO, a.out, H16<E>::TMAP, 4, dims, so, box,
                                    tsw));
    const int64_t dk[3] = {HD, T, (int64_t)H * n_pos};
    const int64_t sk[2] = {HD * 2, (int64_t)T * HD * 2};
    const int bk[3] = {HD, Tk, H};
    GN_PROPAGATE(make_tensor_map_nd(&tmK, kc, H16<E>::TMAP, 3, dk, sk, bk, tsw));
    GN_PROPAGATE(make_tensor_map_nd(&tmV, vc, H16<E>::TMAP, 3, dk, sk, bk, tsw));
  }
  const int rq = (Tq * H * HD * 2 + 1023) & ~1023, rk = (Tk * H * HD * 2 + 1023) & ~1023;
  const int stage = rq + 2 * rk;
  int per_sm = 2, ns = (108 * 1024) / stage;
  if (ns < 3) {
    per_sm = 1;
    ns = (222 * 1024) / stage;
  }
  if (ns > 8) ns = 8;
  GN_REQUIRE(ns >= 2, "temporal attention v2: a stage of %d bytes does not fit twice in shared memory", stage);
  const int smem = ns * stage + 1024 + 2 * ns * 8;
  auto kern = temporal_attn_v2_kernel<E, HD>;
  static DevSmemOptIn optin;
  GN_CUDA_CHECK(ensure_smem_optin(optin, kern, 227 * 1024));
  const int sms = device_sm_count();
  const int grid = std::min(n_pos, sms * per_sm);
  const char* he = getenv("GENIE_B200_KV_HINT");   // 0: plain loads of the K/V cache (A/B measurements)
  const int hint = !(he && he[0] == '0');
  GN_CUDA_CHECK(launch_kernel(PC_TEMPORAL, kern, dim3(grid), dim3((H + 1) * 32), (size_t)smem, st, tmQ, tmK, tmV, tmO,
                              n_pos, S, H, t0, Tq, ns, rq, rk, a.scale * 1.4426950408889634f, hint));
  ++g_launch_count;
  return GN_OK;
}
}  // namespace

bool fast_spatial_supported(const AttnArgs& a, int S) {
  return a.act_bf16 && (a.head_dim == 64 || a.head_dim == 32) && (S == 128 || S == 256);
}
int fast_spatial_attention(const AttnArgs& a, int n_frames, int S, cudaStream_t st) {
  if (a.fp16)
    return a.head_dim == 64 ? launch_spatial_t<64, f16>(a, n_frames, S, st) : launch_spatial_t<32, f16>(a, n_frames, S, st);
  return a.head_dim == 64 ? launch_spatial_t<64, bf16>(a, n_frames, S, st) : launch_spatial_t<32, bf16>(a, n_frames, S, st);
}
bool fast_temporal_supported(const AttnArgs& a, int T) {
  return a.act_bf16 && (a.head_dim == 64 || a.head_dim == 32) && T <= 16 && a.n_heads <= 16;
}
int fast_temporal_attention(const AttnArgs& a, int B, int S, int T, int t0, int Tq, void* kcache, void* vcache,
                            cudaStream_t st) {
  if (a.fp16)
    return a.head_dim == 64 ? launch_temporal_t<64, f16>(a, B, S, T, t0, Tq, kcache, vcache, st)
                            : launch_temporal_t<32, f16>(a, B, S, T, t0, Tq, kcache, vcache, st);
  return a.head_dim == 64 ? launch_temporal_t<64, bf16>(a, B, S, T, t0, Tq, kcache, vcache, st)
                          : launch_temporal_t<32, bf16>(a, B, S, T, t0, Tq, kcache, vcache, st);
}

int generic_temporal_attention(const AttnArgs& a, int B, int S, int T, int t0, int Tq, void* kcache, void* vcache,
                               cudaStream_t st);

bool temporal_v2_supported(const AttnArgs& a, int S, int T) {
  return a.act_bf16 && (a.head_dim == 64 || a.head_dim == 32) && a.n_heads <= 16 && T <= 16 && S % 32 == 0 &&
         a.qk_gamma == nullptr;
}
int launch_temporal_attention_v2(const AttnArgs& a, int nb, int S, int T, int t0, int Tq, const void* kcache,
                                 const void* vcache, cudaStream_t st) {
  GN_REQUIRE(temporal_v2_supported(a, S, T), "temporal attention v2: unsupported shape");
  GN_REQUIRE(t0 >= 0 && Tq >= 1 && t0 + Tq <= T, "temporal attention: bad frame range t0=%d Tq=%d T=%d", t0, Tq, T);
  GN_REQUIRE(kcache && vcache, "temporal attention v2 reads K/V from the head-major caches");
  if (a.head_dim == 32)
    return a.fp16 ? launch_temporal_v2<f16, 32>(a, nb, S, T, t0, Tq, kcache, vcache, st)
                  : launch_temporal_v2<bf16, 32>(a, nb, S, T, t0, Tq, kcache, vcache, st);
  return a.fp16 ? launch_temporal_v2<f16, 64>(a, nb, S, T, t0, Tq, kcache, vcache, st)
                : launch_temporal_v2<bf16, 64>(a, nb, S, T, t0, Tq, kcache, vcache, st);
}

int launch_spatial_attention(const AttnArgs& a, int n_frames, int S, int force_generic, cudaStream_t st) {
  if (!force_generic && tc_spatial_supported(a, S)) return tc_spatial_attention(a, n_frames, S, st);
  if (!force_generic) ++g_fallback_launches;   // a 16-bit handle left the tcgen05 kernel (head_dim != 64, qk-LN in kernel, S)
  if (!force_generic && fast_spatial_supported(a, S)) return fast_spatial_attention(a, n_frames, S, st);
  return launch_generic_attention(a, n_frames, S, 0, st);
}
int launch_temporal_attention(const AttnArgs& a, int B, int S, int T, int t0, int Tq, void* kcache, void* vcache,
                              int force_generic, cudaStream_t st) {
  GN_REQUIRE(t0 >= 0 && Tq >= 1 && t0 + Tq <= T, "temporal attention: bad frame range t0=%d Tq=%d T=%d", t0, Tq, T);
  GN_REQUIRE(t0 == 0 || (kcache && vcache), "temporal attention with t0 > 0 needs the K/V caches");
  if (!force_generic) ++g_fallback_launches;   // a 16-bit handle is not on the TMA-fed v2 kernel
  if (!force_generic && fast_temporal_supported(a, T))
    return fast_temporal_attention(a, B, S, T, t0, Tq, kcache, vcache, st);
  return generic_temporal_attention(a, B, S, T, t0, Tq, kcache, vcache, st);
}

}  // namespace gn
