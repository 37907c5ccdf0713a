// Spatial attention of the ST block (st_transformer.py:73-75, attention.py:36-61) on the 5th-generation tensor
// cores: non-causal softmax(q k^T * scale) v over the S = 256 (or 128) tokens of one frame, head_dim 64, bf16.
//
// One CTA = one 128-query tile of one (frame, head); 2 CTAs are resident per SM (S TMEM columns each), so the
// softmax of one tile overlaps the loads / MMAs of the other.
//   TMA   : Q [128,64], K [S,64], V [S,64] boxes of the fused QKV projection output [rows, 3d] (column order
//           (3, h, hd), attention.py:38) -> 128B-swizzled shared memory; no (B T) S C transpose exists anywhere.
//   MMA 1 : scores[128,S] = Q K^T   tcgen05.mma kind::f16, both operands K-major from smem, fp32 in TMEM cols [0,S)
//   softmax: 256 threads, two per query row (TMEM lane = row; each thread owns half of the S score columns): row max
//           (partials exchanged through smem), p = 2^(s*scale*log2e - max*scale*log2e) in fp32, row sum in fp32,
//           P rounded to bf16 and written back into TMEM over the thread's own, already consumed score columns
//   MMA 2 : O[128,64] = P V         tcgen05.mma with the A operand read from TMEM (no shared-memory round trip for P),
//           V as the MN-major B operand straight from its TMA box, fp32 in TMEM cols [S/4, S/4+64)
//   out   : O / rowsum -> bf16 -> swizzled staging in the (dead) Q tile -> one TMA store into out[rows, d]
// The row sum uses the un-rounded fp32 probabilities and normalisation happens after PV, like the mma.sync kernel
// in attention_fast.cu that this one replaces for head_dim 64 without qk-LayerNorm.
#include "kernels.cuh"
#include "tensormap.cuh"

namespace gn {
namespace {

constexpr int TCA_THREADS = 256;   // 2 column groups x 128 query rows
constexpr int TCA_QROWS = 128;
constexpr int TCA_HD = 64;
constexpr int TCA_ROWB = TCA_HD * 2;   // bytes per staged row = one 128B swizzle atom

// MN-major B operand (V: [keys, 64] row-major = head_dim contiguous), 128-byte swizzle: 8 key rows x 128 B form
// one swizzle atom, 8-row groups are 1024 B apart (SBO); N = 64 is a single 128 B group, so LBO is not used.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int SK, typename H>
__global__ void __launch_bounds__(TCA_THREADS, 2)
spatial_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                       const __grid_constant__ CUtensorMap tmO, int d, float scale_log2e) {
  static_assert(SK == 128 || SK == 256, "keys per frame");
  constexpr uint32_t IDESC_QK = umma_idesc(H16<H>::UMMA_FMT, TCA_QROWS, SK, 0, 0);
  constexpr uint32_t IDESC_PV = umma_idesc(H16<H>::UMMA_FMT, TCA_QROWS, TCA_HD, 0, 1);
  constexpr int GC = SK / 2;      // score columns per thread: two threads (column groups) share a query row
  // group g keeps its packed P in columns [g*GC, g*GC + GC/2).  O takes 64 columns that hold no P: the consumed second
  // half of group 0's scores when that is wide enough (S = 256), else columns past the scores (S = 128)
  constexpr int O_COL = (GC / 2 >= TCA_HD) ? GC / 2 : SK;
  constexpr int TMEM_COLS = 256;
  static_assert(O_COL + TCA_HD <= TMEM_COLS, "O accumulator outside the allocation");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TCA_QROWS * TCA_ROWB;
  uint8_t* sV = sK + SK * TCA_ROWB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + SK * TCA_ROWB);   // [0] Q+K landed [1] V landed [2] scores [3] O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* red = reinterpret_cast<float*>(bars + 6);                    // [2 stats][2 groups][128 rows]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp >> 2;                      // column group
  const int row = (warp & 3) * 32 + lane;         // query row = TMEM lane (a warp reaches lane quarter warp % 4)
  const int qt = blockIdx.x, h = blockIdx.y, f = blockIdx.z;
  const int frame_row0 = f * SK;
  const int q_row0 = frame_row0 + qt * TCA_QROWS;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], (TCA_QROWS + SK) * TCA_ROWB);
    tma_load_2d(sQ, &tmQ, &bars[0], h * TCA_HD, q_row0);
    tma_load_2d(sK, &tmKV, &bars[0], d + h * TCA_HD, frame_row0);
    mbar_arrive_expect_tx(&bars[1], SK * TCA_ROWB);
    tma_load_2d(sV, &tmKV, &bars[1], 2 * d + h * TCA_HD, frame_row0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint64_t qd = umma_desc_kmajor_sw128(smem_u32(sQ));
    const uint64_t kd = umma_desc_kmajor_sw128(smem_u32(sK));
#pragma unroll
    for (int k = 0; k < TCA_HD / 16; ++k) umma_bf16(tmem, qd + 2 * k, kd + 2 * k, IDESC_QK, k != 0);
    umma_commit(&bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);
  tc_fence_after();

  // ---- softmax: this thread owns columns [grp*GC, grp*GC + GC) of query row `row`
  const uint32_t my_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + grp * GC;
  {
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    uint32_t r0[32], r1[32];
    tmem_ld_32x32b_x32(my_addr, r0);
#pragma unroll
    for (int c = 0; c < GC / 32; c += 2) {
      tmem_ld_32x32b_x32(my_addr + (c + 1) * 32, r1);
      tmem_ld_wait();   // waits for both; r0 was requested one iteration earlier
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(r0[i]));
        m1 = fmaxf(m1, __uint_as_float(r0[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(r0[i + 2]));
        m3 = fmaxf(m3, __uint_as_float(r0[i + 3]));
      }
      if (c + 2 < GC / 32) tmem_ld_32x32b_x32(my_addr + (c + 2) * 32, r0);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(r1[i]));
        m1 = fmaxf(m1, __uint_as_float(r1[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(r1[i + 2]));
        m3 = fmaxf(m3, __uint_as_float(r1[i + 3]));
      }
    }
    red[grp * TCA_QROWS + row] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  }
  __syncthreads();
  const float off = fmaxf(red[row], red[TCA_QROWS + row]) * scale_log2e;
  {
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    uint32_t r0[32], r1[32], pk[16];
    tmem_ld_32x32b_x32(my_addr, r0);
#pragma unroll
    for (int c = 0; c < GC / 32; c += 2) {
      tmem_ld_32x32b_x32(my_addr + (c + 1) * 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(r0[2 * i]), scale_log2e, -off));
        const float p1 = ex2_approx(fmaf(__uint_as_float(r0[2 * i + 1]), scale_log2e, -off));
        const float p2 = ex2_approx(fmaf(__uint_as_float(r0[2 * i + 2]), scale_log2e, -off));
        const float p3 = ex2_approx(fmaf(__uint_as_float(r0[2 * i + 3]), scale_log2e, -off));
        l0 += p0; l1 += p1; l2 += p2; l3 += p3;
        pk[i] = pack_h2<H>(p0, p1);
        pk[i + 1] = pack_h2<H>(p2, p3);
      }
      // packed P columns [16c, 16c+16) of this group lie inside its own, already consumed, score columns
      tmem_st_32x32b_x16(my_addr + c * 16, pk);
      if (c + 2 < GC / 32) tmem_ld_32x32b_x32(my_addr + (c + 2) * 32, r0);
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(r1[2 * i]), scale_log2e, -off));
        const float p1 = ex2_approx(fmaf(__uint_as_float(r1[2 * i + 1]), scale_log2e, -off));
        const float p2 = ex2_approx(fmaf(__uint_as_float(r1[2 * i + 2]), scale_log2e, -off));
        const float p3 = ex2_approx(fmaf(__uint_as_float(r1[2 * i + 3]), scale_log2e, -off));
        l0 += p0; l1 += p1; l2 += p2; l3 += p3;
        pk[i] = pack_h2<H>(p0, p1);
        pk[i + 1] = pack_h2<H>(p2, p3);
      }
      tmem_st_32x32b_x16(my_addr + (c + 1) * 16, pk);
    }
    red[(2 + grp) * TCA_QROWS + row] = (l0 + l1) + (l2 + l3);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    const uint64_t vd = umma_desc_mnmajor_sw128(smem_u32(sV));
#pragma unroll
    for (int k = 0; k < SK / 16; ++k) {   // 16 keys per MMA: 8 TMEM columns of packed P, 16 rows (2048 B) of V
      const uint32_t pa = tmem + (k / (GC / 16)) * GC + (k % (GC / 16)) * 8;
      umma_bf16_ts(tmem + O_COL, pa, vd + (uint64_t)(k * (2048 >> 4)), IDESC_PV, k != 0);
    }
    umma_commit(&bars[3]);
  }
  __syncwarp();
  mbar_wait(&bars[3], 0);
  tc_fence_after();

  // ---- normalise (row sum = group 0 + group 1, fixed order), stage the 128 x 64 bf16 tile in the Q buffer (its
  //      MMA has retired): this thread converts head_dim columns [32 grp, 32 grp + 32) of its row; one TMA store
  const float inv = 1.f / (red[2 * TCA_QROWS + row] + red[3 * TCA_QROWS + row]);
  {
    uint32_t r0[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + O_COL + grp * 32, r0);
    tmem_ld_wait();
    uint8_t* srow = sQ + row * TCA_ROWB;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 v;
      v.x = pack_h2<H>(__uint_as_float(r0[8 * c + 0]) * inv, __uint_as_float(r0[8 * c + 1]) * inv);
      v.y = pack_h2<H>(__uint_as_float(r0[8 * c + 2]) * inv, __uint_as_float(r0[8 * c + 3]) * inv);
      v.z = pack_h2<H>(__uint_as_float(r0[8 * c + 4]) * inv, __uint_as_float(r0[8 * c + 5]) * inv);
      v.w = pack_h2<H>(__uint_as_float(r0[8 * c + 6]) * inv, __uint_as_float(r0[8 * c + 7]) * inv);
      *reinterpret_cast<uint4*>(srow + (((grp * 4 + c) ^ (row & 7)) << 4)) = v;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tma_store_2d(&tmO, sQ, h * TCA_HD, q_row0);
    tma_store_commit();
    tma_store_wait_read<0>();   // the staging tile must outlive the read; completion is covered by kernel exit
  }
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

template <int SK, typename H>
int launch_tc_t(const AttnArgs& a, int n_frames, cudaStream_t st) {
  const int d = a.n_heads * a.head_dim;
  const int64_t rows = (int64_t)n_frames * SK;
  CUtensorMap tmQ, tmKV, tmO;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  GN_PROPAGATE(make_tensor_map_2d(&tmQ, a.qkv, H16<H>::TMAP, 2, 3 * d, rows, 3 * d, TCA_HD,
                                  TCA_QROWS, sw));
  GN_PROPAGATE(make_tensor_map_2d(&tmKV, a.qkv, H16<H>::TMAP, 2, 3 * d, rows, 3 * d, TCA_HD, SK, sw));
  GN_PROPAGATE(make_tensor_map_2d(&tmO, a.out, H16<H>::TMAP, 2, d, rows, d, TCA_HD, TCA_QROWS, sw));
  const int smem = (TCA_QROWS + 2 * SK) * TCA_ROWB + 64 + 4 * TCA_QROWS * 4 + 1024;
  auto kern = spatial_attn_tc_kernel<SK, H>;
  static DevSmemOptIn optin;
  GN_CUDA_CHECK(ensure_smem_optin(optin, kern, smem));
  dim3 grid(SK / TCA_QROWS, a.n_heads, n_frames);
  GN_CUDA_CHECK(launch_kernel(PC_SPATIAL, kern, grid, dim3(TCA_THREADS), (size_t)smem, st, tmQ, tmKV, tmO, d,
                              a.scale * 1.4426950408889634f));
  ++g_launch_count;
  return GN_OK;
}

// =====================================================================================
// Persistent, warp-specialised variant for S = 256 (the production shape): one CTA per SM loops over (frame, head)
// items; a 2-stage TMA ring prefetches the next item's Q (both 128-query tiles), K and V while the current one is
// computed, so K/V travel L2 -> SM once per item and no load latency is exposed.
//   warp 0    : TMA producer                       warp 1 : tcgen05.mma issuer (+ TMEM allocation, 512 columns)
//   warps 2-5 : softmax group of query tile 0      warps 6-9 : softmax group of query tile 1
// The kernel is bound by the TMEM read port (64 B/clk per SM = 16 fp32 scores per clock, the same rate as MUFU.EX2),
// so every score is read from TMEM exactly ONCE: a softmax thread owns one query row and handles its 256 keys as two
// blocks of 128 that it holds entirely in registers (block max -> exponentials -> packed bf16 P back into the
// block's own TMEM columns).  The two blocks keep their own maxima m_a, m_b, row sums and PV accumulators O_a, O_b;
// the epilogue merges them exactly:  O = (2^(m_a-m) O_a + 2^(m_b-m) O_b) / (2^(m_a-m) l_a + 2^(m_b-m) l_b),
// m = max(m_a, m_b)  (no rescaling pass over O, no second pass over the scores).
// TMEM, per query tile j (columns relative to 256 j): scores [0,256); block a: P_a [0,64), O_a [64,128); block b:
// P_b [128,192), O_b [192,256).  The issuer orders  PV_a(0) PV_a(1) PV_b(0) PV_b(1)  of item i and interleaves the
// next item's QK as soon as a tile's accumulators have been read, so one group's exponentials overlap the other
// group's TMEM loads and the MMAs.  Output tiles are staged in the (dead) Q tile of the item's stage and written with
// one TMA store per tile; the stage returns to the producer when its MMAs have retired and both stores have read it.
// =====================================================================================
constexpr int TCP_THREADS = 32 * 10;
constexpr int TCP_SK = 256;
constexpr int TCP_STAGE_BYTES = 3 * TCP_SK * TCA_ROWB;   // Q (2 tiles) + K + V = 96 KB
constexpr int TCP_STAGES = 2;

struct TcpBars {
  uint64_t full[TCP_STAGES], empty[TCP_STAGES];
  uint64_t sready[2], pready[2][2], oready[2], tfree[2];
  uint32_t tmem_slot, pad;
};

// HD = 32 (the in-tree 35M config, genie/configs/magvit_n32_h8_d256.json): the TMA boxes stay 64 columns wide and cover
// the head PAIR (h & ~1, h | 1) - one 128-byte swizzle atom per row, exactly the HD = 64 shared-memory layout.  Head h
// uses the 32-byte K slices [2 (h & 1), 2 (h & 1) + 2) of the atom for Q K^T (two MMAs instead of four); P V runs over
// the full 64-wide V box (N = 64: the other head's 32 columns are computed and ignored - 4 S d FLOPs of a kernel that is
// bound by the exponentials, not by the tensor pipe), and the epilogue reads, merges and stores only the head's own 32
// accumulator columns (64-byte rows, SWIZZLE_64B staging and store box).
template <typename H, int HD>
__global__ void __launch_bounds__(TCP_THREADS, 1)
spatial_attn_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmO,
                                  int d, int n_heads, int n_items, float scale_log2e) {
  static_assert(HD == 64 || HD == 32, "head_dim 64 or 32");
  constexpr int SK = TCP_SK;
  constexpr int BK = SK / 2;          // keys per softmax block
  constexpr uint32_t IDESC_QK = umma_idesc(H16<H>::UMMA_FMT, TCA_QROWS, SK, 0, 0);
  constexpr uint32_t IDESC_PV = umma_idesc(H16<H>::UMMA_FMT, TCA_QROWS, TCA_HD, 0, 1);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // stage s: [Q tile 0 | Q tile 1 | K | V], each sub-tile 1024-aligned
  TcpBars* bars = reinterpret_cast<TcpBars*>(smem + TCP_STAGES * TCP_STAGE_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_my = (int)blockIdx.x < n_items ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < TCP_STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 3);      // MMAs of the item retired (tcgen05.commit) + the two output stores read
    }
    for (int j = 0; j < 2; ++j) {
      mbar_init(&bars->sready[j], 1);
      mbar_init(&bars->pready[j][0], 4);  // one arrival per softmax warp of the group
      mbar_init(&bars->pready[j][1], 4);
      mbar_init(&bars->oready[j], 1);
      mbar_init(&bars->tfree[j], 4);
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc<512>(&bars->tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      for (int i = 0; i < n_my; ++i) {
        const int item = (int)blockIdx.x + i * (int)gridDim.x;
        const int f = item / n_heads, h = item % n_heads;
        const int s = i & 1;
        uint8_t* st = smem + s * TCP_STAGE_BYTES;
        mbar_wait(&bars->empty[s], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->full[s], TCP_STAGE_BYTES);
        // the fused QKV rows are read exactly once (by this kernel; twice for HD = 32, from L2): evict_first
        const int c0 = HD == 64 ? h * TCA_HD : (h >> 1) * TCA_HD;       // first column of the 64-wide box
        tma_load_2d_hint(st, &tmKV, &bars->full[s], c0, f * SK, pol);                              // Q, 256 rows
        tma_load_2d_hint(st + SK * TCA_ROWB, &tmKV, &bars->full[s], d + c0, f * SK, pol);          // K
        tma_load_2d_hint(st + 2 * SK * TCA_ROWB, &tmKV, &bars->full[s], 2 * d + c0, f * SK, pol);  // V
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && n_my > 0) {
      auto issue_qk = [&](int i, int j) {   // scores of query tile j of item i
        const uint8_t* st = smem + (i & 1) * TCP_STAGE_BYTES;
        const uint64_t qd = umma_desc_kmajor_sw128(smem_u32(st + j * TCA_QROWS * TCA_ROWB));
        const uint64_t kd = umma_desc_kmajor_sw128(smem_u32(st + SK * TCA_ROWB));
        // HD = 32: this head's two 32-byte K slices inside the pair's 128-byte atom
        const int k0 = HD == 64 ? 0 : 2 * ((((int)blockIdx.x + i * (int)gridDim.x) % n_heads) & 1);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem + j * SK, qd + 2 * (k0 + k), kd + 2 * (k0 + k), IDESC_QK, k != 0);
        umma_commit(&bars->sready[j]);
      };
      auto issue_pv = [&](int i, int j, int blk) {   // O_blk = P_blk V[128 blk .. 128 blk + 128)
        const uint8_t* st = smem + (i & 1) * TCP_STAGE_BYTES;
        const uint64_t vd = umma_desc_mnmajor_sw128(smem_u32(st + 2 * SK * TCA_ROWB + blk * BK * TCA_ROWB));
        const uint32_t tb = tmem + j * SK + blk * BK;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)   // 16 keys per MMA: 8 TMEM columns of packed P, 16 rows (2048 B) of V
          umma_bf16_ts(tb + BK / 2, tb + k * 8, vd + (uint64_t)(k * (2048 >> 4)), IDESC_PV, k != 0);
      };
      // The two groups are started one softmax block apart (QK of tile 1 of the first item is held back until tile
      // 0's first block is done) and stay staggered: while one group runs its exponentials (MUFU), the other one
      // loads its next 128 scores (TMEM read port), instead of both contending for the same unit in lockstep.
      mbar_wait(&bars->full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      for (int i = 0; i < n_my; ++i) {
        const uint32_t ph = i & 1;
        const bool more = i + 1 < n_my;
        mbar_wait(&bars->pready[0][0], ph);
        tc_fence_after();
        issue_pv(i, 0, 0);
        if (i == 0) issue_qk(0, 1);
        mbar_wait(&bars->pready[1][0], ph);
        tc_fence_after();
        issue_pv(i, 1, 0);
        mbar_wait(&bars->pready[0][1], ph);
        tc_fence_after();
        issue_pv(i, 0, 1);
        umma_commit(&bars->oready[0]);
        mbar_wait(&bars->pready[1][1], ph);
        tc_fence_after();
        issue_pv(i, 1, 1);
        umma_commit(&bars->oready[1]);
        umma_commit(&bars->empty[i & 1]);      // every MMA reading stage (i & 1) has been issued
        if (more) {
          mbar_wait(&bars->full[(i + 1) & 1], ((i + 1) >> 1) & 1);
          mbar_wait(&bars->tfree[0], ph);      // group 0 has read O(i): tile 0's columns may be overwritten
          tc_fence_after();
          issue_qk(i + 1, 0);
          mbar_wait(&bars->tfree[1], ph);
          tc_fence_after();
          issue_qk(i + 1, 1);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax groups: one thread per query row
    const int j = (warp - 2) >> 2;                 // query tile of this group
    const int q4 = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q4 * 32 + lane;
    const int gtid = ((warp - 2) & 3) * 32 + lane; // 0..127 within the group
    const uint32_t my_addr = tmem + ((uint32_t)(q4 * 32) << 16) + j * SK;
    for (int i = 0; i < n_my; ++i) {
      const uint32_t ph = i & 1;
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int f = item / n_heads, h = item % n_heads;
      mbar_wait(&bars->sready[j], ph);
      tc_fence_after();
      float mblk[2], lblk[2];
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        uint32_t r[BK];
#pragma unroll
        for (int c = 0; c < BK / 32; ++c)
          tmem_ld_32x32b_x32(my_addr + blk * BK + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
        tmem_ld_wait();
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int e = 0; e < BK; e += 4) {
          m0 = fmaxf(m0, __uint_as_float(r[e]));
          m1 = fmaxf(m1, __uint_as_float(r[e + 1]));
          m2 = fmaxf(m2, __uint_as_float(r[e + 2]));
          m3 = fmaxf(m3, __uint_as_float(r[e + 3]));
        }
        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        const float off = mx * scale_log2e;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
        for (int c = 0; c < BK / 32; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(r[c * 32 + 2 * e]), scale_log2e, -off));
            const float p1 = ex2_approx(fmaf(__uint_as_float(r[c * 32 + 2 * e + 1]), scale_log2e, -off));
            const float p2 = ex2_approx(fmaf(__uint_as_float(r[c * 32 + 2 * e + 2]), scale_log2e, -off));
            const float p3 = ex2_approx(fmaf(__uint_as_float(r[c * 32 + 2 * e + 3]), scale_log2e, -off));
            l0 += p0; l1 += p1; l2 += p2; l3 += p3;
            pk[e] = pack_h2<H>(p0, p1);
            pk[e + 1] = pack_h2<H>(p2, p3);
          }
          tmem_st_32x32b_x16(my_addr + blk * BK + c * 16, pk);   // over the block's own, already loaded scores
        }
        mblk[blk] = mx;
        lblk[blk] = (l0 + l1) + (l2 + l3);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->pready[j][blk]);
      }

      mbar_wait(&bars->oready[j], ph);
      tc_fence_after();
      uint8_t* stq = smem + (i & 1) * TCP_STAGE_BYTES + j * TCA_QROWS * TCA_ROWB;   // dead Q tile = output staging
      {
        uint32_t ra[HD], rb[HD];
        const int oc = HD == 64 ? 0 : 32 * (h & 1);      // this head's accumulator columns inside the 64-wide O
        tmem_ld_32x32b_x32(my_addr + BK / 2 + oc, *reinterpret_cast<uint32_t(*)[32]>(&ra[0]));
        if (HD == 64) tmem_ld_32x32b_x32(my_addr + BK / 2 + 32, *reinterpret_cast<uint32_t(*)[32]>(&ra[HD - 32]));
        tmem_ld_32x32b_x32(my_addr + BK + BK / 2 + oc, *reinterpret_cast<uint32_t(*)[32]>(&rb[0]));
        if (HD == 64) tmem_ld_32x32b_x32(my_addr + BK + BK / 2 + 32, *reinterpret_cast<uint32_t(*)[32]>(&rb[HD - 32]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tfree[j]);   // tile j's TMEM columns are free for QK(i+1)
        const float m = fmaxf(mblk[0], mblk[1]);
        float wa = ex2_approx((mblk[0] - m) * scale_log2e);
        float wb = ex2_approx((mblk[1] - m) * scale_log2e);
        const float inv = 1.f / (wa * lblk[0] + wb * lblk[1]);
        wa *= inv;
        wb *= inv;
        uint8_t* srow = stq + row * (HD * 2);          // 128-byte rows / SWIZZLE_128B, or 64-byte rows / SWIZZLE_64B
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o[e] = fmaf(__uint_as_float(ra[8 * c + e]), wa, __uint_as_float(rb[8 * c + e]) * wb);
          uint4 v;
          v.x = pack_h2<H>(o[0], o[1]);
          v.y = pack_h2<H>(o[2], o[3]);
          v.z = pack_h2<H>(o[4], o[5]);
          v.w = pack_h2<H>(o[6], o[7]);
          const int sc = HD == 64 ? (c ^ (row & 7)) : (c ^ ((row >> 1) & 3));
          *reinterpret_cast<uint4*>(srow + (sc << 4)) = v;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + j, 128);
      if (gtid == 0) {
        tma_store_2d(&tmO, stq, h * HD, f * SK + j * TCA_QROWS);
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&bars->empty[i & 1]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <typename H, int HD>
int launch_tc_persistent(const AttnArgs& a, int n_frames, cudaStream_t st) {
  const int d = a.n_heads * a.head_dim;
  const int64_t rows = (int64_t)n_frames * TCP_SK;
  CUtensorMap tmKV, tmO;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  GN_PROPAGATE(make_tensor_map_2d(&tmKV, a.qkv, H16<H>::TMAP, 2, 3 * d, rows, 3 * d, TCA_HD, TCP_SK,
                                  sw));
  GN_PROPAGATE(make_tensor_map_2d(&tmO, a.out, H16<H>::TMAP, 2, d, rows, d, HD, TCA_QROWS,
                                  HD == 64 ? sw : CU_TENSOR_MAP_SWIZZLE_64B));
  const int smem = TCP_STAGES * TCP_STAGE_BYTES + (int)sizeof(TcpBars) + 1024;
  static DevSmemOptIn optin;
  GN_CUDA_CHECK(ensure_smem_optin(optin, spatial_attn_tc_persistent_kernel<H, HD>, smem));
  const int sms = device_sm_count();
  const int n_items = n_frames * a.n_heads;
  const int grid = n_items < sms ? n_items : sms;
  GN_CUDA_CHECK(launch_kernel(PC_SPATIAL, spatial_attn_tc_persistent_kernel<H, HD>, dim3(grid), dim3(TCP_THREADS),
                              (size_t)smem, st, tmKV, tmO, d, a.n_heads, n_items, a.scale * 1.4426950408889634f));
  ++g_launch_count;
  return GN_OK;
}

}  // namespace

bool tc_spatial_supported(const AttnArgs& a, int S) {
  const char* e = getenv("GENIE_B200_SPATIAL_TC");   // re-read per launch (tests compare both kernels in one process)
  const bool on = !(e && (e[0] == '0' || e[0] == 'n' || e[0] == 'N'));
  if (!on || !a.act_bf16 || a.qk_gamma != nullptr) return false;
  if (a.head_dim == TCA_HD) return S == 128 || S == 256;
  // head_dim 32: persistent kernel only (S = 256), heads handled through 64-wide head-pair boxes
  return a.head_dim == 32 && S == 256 && a.n_heads % 2 == 0 && !(e && e[0] == '1');
}

int tc_spatial_attention(const AttnArgs& a, int n_frames, int S, cudaStream_t st) {
  // GENIE_B200_SPATIAL_TC: unset / 2 = persistent kernel for S = 256, 1 = one-CTA-per-tile kernel, 0 = mma.sync kernel
  const char* e = getenv("GENIE_B200_SPATIAL_TC");
  const bool persistent = !(e && e[0] == '1');
  if (a.head_dim == 32)
    return a.fp16 ? launch_tc_persistent<f16, 32>(a, n_frames, st) : launch_tc_persistent<bf16, 32>(a, n_frames, st);
  if (a.fp16) {
    if (S == 256)
      return persistent ? launch_tc_persistent<f16, 64>(a, n_frames, st) : launch_tc_t<256, f16>(a, n_frames, st);
    return launch_tc_t<128, f16>(a, n_frames, st);
  }
  if (S == 256)
    return persistent ? launch_tc_persistent<bf16, 64>(a, n_frames, st) : launch_tc_t<256, bf16>(a, n_frames, st);
  return launch_tc_t<128, bf16>(a, n_frames, st);
}

}  // namespace gn
