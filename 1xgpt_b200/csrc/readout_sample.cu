// Readout fused with the MaskGIT decode math ("factored softmax / argmax: logits never hit HBM").
//   reference: genie/st_mask_git.py:262 (out_x_proj) + :171-190 (per factored vocabulary: softmax over 512 logits,
//   argmax, confidence = prod_i p_i[sample_i], id = sum_i sample_i * 512^i with the high vocabulary first)
//
// One CTA = 128 token rows of the frame being decoded.  For each factored vocabulary (high one first) the CTA computes
// the row's 512 logits  x[128, d] . W_v[512, d]^T  with tcgen05.mma (two N = 256 instructions per 32-byte K slice, fp32
// accumulators filling all 512 TMEM columns), operands staged by TMA through a 2-stage ring (A k-block 16 KB + the
// vocabulary's weight k-block 64 KB).  The epilogue warps own one row per thread (TMEM lane = row): pass 1 adds the
// bias and finds max / first argmax over the row's 512 columns, pass 2 sums exp(l - max); the sample id and the
// confidence are the only mandatory outputs (8 bytes per token instead of 4 KB of logits).  When the caller needs the
// step-0 logits (maskgit_generate's return value, evaluate's CE) they are staged and TMA-stored from pass 1.
// temperature > 0 (Categorical draw) keeps the two-kernel path: readout GEMM + sample_kernel (decode.cu).
#include "kernels.cuh"
#include "tensormap.cuh"
#include <cfloat>

namespace gn {
namespace {

constexpr int RS_ROWS = 128;
constexpr int RS_V = 512;                      // logits per factored vocabulary = TMEM columns
constexpr int RS_A_BYTES = RS_ROWS * 128;      // one 64-element K block of A
constexpr int RS_B_BYTES = RS_V * 128;         // ... of the vocabulary's weights (two TMA boxes of 256 rows)
constexpr int RS_STAGE_BYTES = RS_A_BYTES + RS_B_BYTES;
constexpr int RS_STAGES = 2;
constexpr int RS_EPI_WARPS = 4;
constexpr int RS_THREADS = 32 * (2 + RS_EPI_WARPS);
constexpr int RS_STAGING = RS_EPI_WARPS * 2 * 4096;   // two 32 x 32 fp32 chunk buffers per warp (logits output only)

struct RsArgs {
  int R, K, NV;
  const float* bias;     // [NV * 512]
  int32_t* samples;      // [R]
  float* conf;           // [R]
  int write_logits;      // 1: also store the logits rows [R, NV * 512] through tmOut
};

template <typename H>
__global__ void __launch_bounds__(RS_THREADS, 1)
readout_sample_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmOut, const RsArgs a) {
  constexpr uint32_t IDESC = umma_idesc(H16<H>::UMMA_FMT, RS_ROWS, 256);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + RS_STAGES * RS_STAGE_BYTES;
  float* sbias = reinterpret_cast<float*>(staging + RS_STAGING);             // [512]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + RS_V);
  uint64_t* full_bar = bars;                  // [RS_STAGES]
  uint64_t* empty_bar = bars + RS_STAGES;     // [RS_STAGES]
  uint64_t* acc_full = bars + 2 * RS_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (int)blockIdx.x * RS_ROWS;
  const int num_kb = (a.K + 63) / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (a.write_logits) tma_prefetch_desc(&tmOut);
    for (int s = 0; s < RS_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, RS_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int vi = 0; vi < a.NV; ++vi) {
        const int v = a.NV - 1 - vi;          // high vocabulary first (st_mask_git.py:179 iterates flip(2))
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * RS_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], RS_STAGE_BYTES);
          tma_load_2d(st, &tmA, &full_bar[stage], kb * 64, row0);
          tma_load_2d(st + RS_A_BYTES, &tmB, &full_bar[stage], kb * 64, v * RS_V);
          tma_load_2d(st + RS_A_BYTES + RS_B_BYTES / 2, &tmB, &full_bar[stage], kb * 64, v * RS_V + 256);
          if (++stage == RS_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int vi = 0; vi < a.NV; ++vi) {
        mbar_wait(acc_empty, (vi & 1) ^ 1);   // the epilogue has read the previous vocabulary's logits
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint8_t* st = smem + stage * RS_STAGE_BYTES;
          const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(st));
          const uint64_t b0 = umma_desc_kmajor_sw128(smem_u32(st + RS_A_BYTES));
          const uint64_t b1 = umma_desc_kmajor_sw128(smem_u32(st + RS_A_BYTES + RS_B_BYTES / 2));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(tmem, adesc + 2 * k, b0 + 2 * k, IDESC, (kb | k) != 0);
            umma_bf16(tmem + 256, adesc + 2 * k, b1 + 2 * k, IDESC, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == RS_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: one row per thread
    const uint32_t q = warp & 3;               // TMEM lane quarter of this warp
    const uint32_t ew = warp - 2;
    const int et = (int)threadIdx.x - 64;      // 0..127 among the epilogue threads
    const int row = row0 + (int)(q * 32 + lane);
    uint8_t* st0 = staging + ew * 2 * 4096;
    const uint32_t acc_addr = tmem + ((q * 32u) << 16);
    int id = 0;
    float cf = 1.f;
    uint32_t buf = 0;
    for (int vi = 0; vi < a.NV; ++vi) {
      const int v = a.NV - 1 - vi;
      named_bar_sync(1, 32 * RS_EPI_WARPS);    // every epilogue thread is done with the previous vocabulary's bias
      *reinterpret_cast<float4*>(sbias + 4 * et) = __ldg(reinterpret_cast<const float4*>(a.bias + v * RS_V) + et);
      named_bar_sync(1, 32 * RS_EPI_WARPS);
      mbar_wait(acc_full, vi & 1);
      tc_fence_after();
      // pass 1: l = acc + bias; max / first argmax; optional logits store
      float mx = -FLT_MAX;
      int arg = 0;
      uint32_t rr[2][32];
      tmem_ld_32x32b_x32(acc_addr, rr[0]);
      // one 32-column chunk: wait for its TMEM load, start the next chunk's load, then the math on this one
      auto pass1 = [&](const uint32_t(&r)[32], uint32_t(&rnext)[32], int c) {
        tmem_ld_wait();
        if (c + 1 < RS_V / 32) tmem_ld_32x32b_x32(acc_addr + (c + 1) * 32, rnext);
        float l[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          l[j] = __uint_as_float(r[j]) + sbias[c * 32 + j];
          if (l[j] > mx) { mx = l[j]; arg = c * 32 + j; }
        }
        if (a.write_logits) {
          if (lane == 0) tma_store_wait_read<1>();      // the buffer written now was stored two chunks ago
          __syncwarp();
          uint8_t* rowp = st0 + buf * 4096 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) = make_float4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmOut, st0 + buf * 4096, v * RS_V + c * 32, row0 + (int)q * 32);
            tma_store_commit();
          }
          buf ^= 1;
        }
      };
#pragma unroll 1
      for (int c = 0; c < RS_V / 32; c += 2) {
        pass1(rr[0], rr[1], c);
        pass1(rr[1], rr[0], c + 1);
      }
      // pass 2: sum of exp(l - max) over the row (the bias is added again: the logits were not kept)
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      tmem_ld_32x32b_x32(acc_addr, rr[0]);
      auto pass2 = [&](const uint32_t(&r)[32], uint32_t(&rnext)[32], int c) {
        tmem_ld_wait();
        if (c + 1 < RS_V / 32) tmem_ld_32x32b_x32(acc_addr + (c + 1) * 32, rnext);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          s0 += __expf(__uint_as_float(r[j]) + sbias[c * 32 + j] - mx);
          s1 += __expf(__uint_as_float(r[j + 1]) + sbias[c * 32 + j + 1] - mx);
          s2 += __expf(__uint_as_float(r[j + 2]) + sbias[c * 32 + j + 2] - mx);
          s3 += __expf(__uint_as_float(r[j + 3]) + sbias[c * 32 + j + 3] - mx);
        }
      };
#pragma unroll 1
      for (int c = 0; c < RS_V / 32; c += 2) {
        pass2(rr[0], rr[1], c);
        pass2(rr[1], rr[0], c + 1);
      }
      // the accumulator columns are free for the next vocabulary's MMAs
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      id = id * RS_V + arg;
      cf *= 1.f / ((s0 + s1) + (s2 + s3));     // p[argmax] = exp(0) / sum
    }
    if (row < a.R) {
      a.samples[row] = id;
      a.conf[row] = cf;
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <typename H>
int launch_rs(const void* A, const void* W, const float* bias, float* logits, int32_t* samples, float* conf, int R, int K,
              int NV, cudaStream_t st) {
  CUtensorMap tmA, tmB, tmO;
  GN_PROPAGATE(make_tensor_map_2d(&tmA, A, H16<H>::TMAP, 2, K, R, K, 64, RS_ROWS, CU_TENSOR_MAP_SWIZZLE_128B));
  GN_PROPAGATE(make_tensor_map_2d(&tmB, W, H16<H>::TMAP, 2, K, (int64_t)NV * RS_V, K, 64, 256, CU_TENSOR_MAP_SWIZZLE_128B));
  if (logits)
    GN_PROPAGATE(make_tensor_map_2d(&tmO, logits, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (int64_t)NV * RS_V, R,
                                    (int64_t)NV * RS_V, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  else
    tmO = tmA;
  const int smem = RS_STAGES * RS_STAGE_BYTES + RS_STAGING + RS_V * 4 + 64 + 1024;
  auto kern = readout_sample_kernel<H>;
  static DevSmemOptIn optin;
  GN_CUDA_CHECK(ensure_smem_optin(optin, kern, smem));
  RsArgs a{R, K, NV, bias, samples, conf, logits ? 1 : 0};
  GN_CUDA_CHECK(launch_kernel(PC_GEMM_STORE, kern, dim3(ceil_div(R, RS_ROWS)), dim3(RS_THREADS), (size_t)smem, st, tmA,
                              tmB, tmO, a));
  g_gemm_flops_issued += 2.0 * R * (double)NV * RS_V * K;
  ++g_launch_count;
  return GN_OK;
}

}  // namespace

bool readout_sample_supported(int V, int K, int fp16_or_bf16) {
  return fp16_or_bf16 && V == RS_V && K % 8 == 0;
}

int launch_readout_sample(const void* A, const void* W, const float* bias, float* logits, int32_t* samples, float* conf,
                          int R, int K, int NV, int fp16, cudaStream_t st) {
  GN_REQUIRE(A && W && bias && samples && conf && R > 0 && NV >= 1, "readout_sample: invalid argument");
  GN_REQUIRE(reinterpret_cast<uintptr_t>(bias) % 16 == 0, "readout_sample: bias must be 16-byte aligned");
  return fp16 ? launch_rs<f16>(A, W, bias, logits, samples, conf, R, K, NV, st)
              : launch_rs<bf16>(A, W, bias, logits, samples, conf, R, K, NV, st);
}

}  // namespace gn
