// See gemm.cuh.  sm_100a only.
#include "gemm.cuh"
#include "ptx.cuh"
#include "tensormap.cuh"
#include <cstdlib>
#include <vector>

namespace gn {

unsigned long long g_launch_count = 0;
// launches that left the intended Blackwell kernel for a slower variant because of a shape / alignment / precision
// cliff (CUDA-core GEMM without force_simt, generic or mma.sync attention where the tcgen05 / TMA kernel was expected).
// Tests assert that it stays 0 on the production shapes (gn_fallback_launches()).
unsigned long long g_fallback_launches = 0;

static bool env_on(const char* name, bool dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  return !(v[0] == '0' || v[0] == 'n' || v[0] == 'N' || v[0] == 'f' || v[0] == 'F');
}
// GENIE_B200_PAIR=0 falls back to single-CTA 128 x 256 tiles (A/B switch for profiling)
static bool g_use_pair = env_on("GENIE_B200_PAIR", true);

double g_gemm_flops_issued = 0.0;

// =====================================================================================
// tcgen05 path
// =====================================================================================
namespace {

constexpr int BLOCK_M = 128;
constexpr int TILE_K_BYTES = 128;            // one 128B swizzle atom along K per stage
constexpr int A_TILE_BYTES = BLOCK_M * TILE_K_BYTES;
constexpr int NUM_ACC_STAGES = 2;
constexpr int EPI_COLS = 32;                 // accumulator columns handled per epilogue step
// Epilogue warps: 4 (one per TMEM lane quarter) for the residual epilogue; 8 (two per quarter, each taking half of the
// tile's columns) for the store / GELU epilogues, whose per-chunk dependency chain (tcgen05.ld -> math -> staging ->
// TMA store) is latency bound with a single warp per SM sub-partition.
template <int EPI> struct EpiCfg { static constexpr int WARPS = EPI == EPI_RESID ? 4 : 8; };
constexpr int MAX_EPI_WARPS = 8;
constexpr int MAX_GEMM_THREADS = 32 * (2 + MAX_EPI_WARPS);
constexpr int SMEM_LIMIT = 232448;           // 227 KB opt-in dynamic shared memory per CTA

// Residual epilogue: in-place staging buffers per warp (4 KB each); RB - 2 residual chunks are requested ahead of use.
// 4 buffers (3 for 256-wide tiles that also emit the bf16 copy, so that a 3-stage operand ring still fits).  A deeper
// residual ring (6 / 5 buffers) was measured SLOWER (proj 50.8 -> 56.5 us, fc2 83 -> 91 us at M = 32768): it costs one
// operand stage, and the operand ring is what keeps the L2 -> SM stream busy.
constexpr int RES_BUFS = 4;
constexpr int res_bufs(int block_n, bool dual, int /*ctas*/) { return (block_n > 128 && dual) ? 3 : RES_BUFS; }

// CTAS = 2: CTA-pair tiles (tcgen05 cta_group::2).  One tile is 256 rows x BLOCK_N columns; each CTA of the pair stages
// its own 128 rows of A and ONE HALF of the B tile, so a k-block costs A + B/2 bytes of L2->SM traffic per SM instead
// of A + B: the linear layers of this model (K = 512 / 2048) are bound by the chip-wide L2->SM throughput
// (~6.2 KB per L2 clock, profiles/), not by the tensor pipe.
template <int BLOCK_N, int EPI, typename OutT, bool DUAL, int CTAS = 1, bool NARROW = false>
struct GemmSmem {
  static constexpr int NUM_EPI_WARPS = EpiCfg<EPI>::WARPS;
  static constexpr int COLS_PER_WARP = BLOCK_N / (NUM_EPI_WARPS / 4);   // columns of the tile one warp handles
  // store / GELU epilogues: one staging buffer per 32-column chunk of the warp's slice (8 KB per warp), so a buffer is
  // rewritten a whole tile after its TMA store was issued and the store's read latency never stalls the warp
  // bf16 outputs of the store / GELU epilogues: two 32-column chunks are staged side by side (128 B per row = one
  // 128B-swizzle atom) and leave as ONE TMA store of 32 rows x 64 columns: full 128-byte lines instead of 64-byte
  // halves (for the temporal K/V caches: one (position, head, frame) row of head_dim 64 per line) and half as many
  // bulk stores per tile
  // NARROW: 32-column staging even for 16-bit outputs (K/V-cache rows of head_dim 32 are 64-byte lines)
  static constexpr bool WIDE = EPI != EPI_RESID && sizeof(OutT) == 2 && COLS_PER_WARP % (2 * EPI_COLS) == 0 && !NARROW;
  static constexpr int EPI_BUF_BYTES = 32 * EPI_COLS * (int)sizeof(OutT) * (WIDE ? 2 : 1);
  static constexpr int CHUNKS_PER_BUF = WIDE ? 2 : 1;
  static constexpr int STORE_BUFS_RAW = (COLS_PER_WARP / (EPI_COLS * CHUNKS_PER_BUF)) * EPI_BUF_BYTES <= 8192
                                            ? COLS_PER_WARP / (EPI_COLS * CHUNKS_PER_BUF) : 8192 / EPI_BUF_BYTES;
  static constexpr int STORE_BUFS = (WIDE && STORE_BUFS_RAW < 2) ? 2 : STORE_BUFS_RAW;
  static constexpr int OUT_BUFS = EPI == EPI_RESID ? res_bufs(BLOCK_N, DUAL, CTAS) : STORE_BUFS;
  static constexpr int B_TILE_BYTES = (BLOCK_N / CTAS) * TILE_K_BYTES;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int OUT_STAGE_BYTES = EPI == EPI_RESID ? 32 * EPI_COLS * (int)sizeof(OutT) : EPI_BUF_BYTES;  // per warp per buffer
  static constexpr int OUT2_STAGE_BYTES = DUAL ? 32 * EPI_COLS * 2 : 0;
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * (OUT_BUFS * OUT_STAGE_BYTES + 2 * OUT2_STAGE_BYTES);
  static constexpr int BIAS_BYTES = 2 * NUM_EPI_WARPS * COLS_PER_WARP * 4;   // per-warp copies of its slice of the
                                                                             // bias and folded-LN column sums
  static constexpr int BAR_BYTES = 1024;
  static constexpr int ALIGN_SLACK = 1024;
  static constexpr int RAW_STAGES = (SMEM_LIMIT - STAGING_BYTES - BIAS_BYTES - BAR_BYTES - ALIGN_SLACK) / STAGE_BYTES;
  static constexpr int STAGES = RAW_STAGES > 8 ? 8 : RAW_STAGES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + STAGING_BYTES + BIAS_BYTES + BAR_BYTES + ALIGN_SLACK;
  static_assert(STAGES >= 3, "pipeline too shallow");
};

struct TcArgs {
  int M, N, K;
  const float* bias;
  const float* resid;
  int64_t ldr;
  int round_tf32;
  // implicit-GEMM convolution (cin_blocks == 0: plain GEMM)
  int cin_blocks, Ho, Wo, tw, th, stride;
  // folded LayerNorm (consumer side) / row statistics (producer side)
  const float* ln_stats;
  int ln_np, ln_d;
  const float* ln_colsum;
  float* stats_out;
  // K/V columns of the temporal QKV projection go to the caches (kv_d == 0: off)
  int kv_d, kv_hd, kv_S, kv_Tact, kv_t0;
  // profiling aid (GENIE_B200_GEMM_DEBUG; results are garbage): bit 0 = store/GELU epilogues release the accumulator
  // without reading or storing it, bit 1 = the producer signals the stages without loading them
  int dbg;
  int a_hint;   // 1: the A operand is dead after this GEMM -> load it with the L2 evict_first policy
  int w_prefetch;   // 1: pull this CTA's first weight tiles into L2 before griddepcontrol.wait
  const float* qkn_g;   // qk-LayerNorm over 64-column groups of output columns [0, qkn_cols) (wide bf16 store epilogue)
  const float* qkn_b;
  int qkn_cols;
  int red_add;   // fp32 store epilogue: out += tile (TMA reduce-add) instead of out = tile
};

template <typename OutT>
__device__ __forceinline__ void stage_row_chunk(uint8_t* buf, uint32_t lane, const float (&v)[EPI_COLS]);

// fp32: 32 cols = 128 B per row, SWIZZLE_128B (16B chunk index ^= row & 7)
template <>
__device__ __forceinline__ void stage_row_chunk<float>(uint8_t* buf, uint32_t lane, const float (&v)[EPI_COLS]) {
  uint8_t* row = buf + lane * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 f = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    *reinterpret_cast<float4*>(row + ((j ^ (lane & 7)) << 4)) = f;
  }
}
// 16-bit (bf16 / fp16): 32 cols = 64 B per row, SWIZZLE_64B (16B chunk index ^= (row >> 1) & 3)
template <typename H>
__device__ __forceinline__ void stage_row_chunk16(uint8_t* buf, uint32_t lane, const float (&v)[EPI_COLS]) {
  uint8_t* row = buf + lane * 64;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 p;
    p.x = pack_h2<H>(v[8 * j + 0], v[8 * j + 1]);
    p.y = pack_h2<H>(v[8 * j + 2], v[8 * j + 3]);
    p.z = pack_h2<H>(v[8 * j + 4], v[8 * j + 5]);
    p.w = pack_h2<H>(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(row + ((j ^ ((lane >> 1) & 3)) << 4)) = p;
  }
}
template <>
__device__ __forceinline__ void stage_row_chunk<bf16>(uint8_t* buf, uint32_t lane, const float (&v)[EPI_COLS]) {
  stage_row_chunk16<bf16>(buf, lane, v);
}
template <>
__device__ __forceinline__ void stage_row_chunk<f16>(uint8_t* buf, uint32_t lane, const float (&v)[EPI_COLS]) {
  stage_row_chunk16<f16>(buf, lane, v);
}

// 16-bit, wide staging: 64 cols = 128 B per row, SWIZZLE_128B; `half` selects the left / right 32 columns
template <typename H>
__device__ __forceinline__ void stage_row_chunk_wide(uint8_t* buf, uint32_t lane, int half, const float (&v)[EPI_COLS]) {
  uint8_t* row = buf + lane * 128;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 p;
    p.x = pack_h2<H>(v[8 * j + 0], v[8 * j + 1]);
    p.y = pack_h2<H>(v[8 * j + 2], v[8 * j + 3]);
    p.z = pack_h2<H>(v[8 * j + 4], v[8 * j + 5]);
    p.w = pack_h2<H>(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(row + (((half * 4 + j) ^ (lane & 7)) << 4)) = p;
  }
}
// the 16-bit type that goes with an operand type: itself for bf16 / fp16, bf16 for fp32 (tf32) operands
template <typename InT> struct Half16Of { typedef InT type; };
template <> struct Half16Of<float> { typedef bf16 type; };

template <typename InT, int BLOCK_N, int EPI, typename OutT, bool DUAL, int CTAS, bool QKN, bool NARROW>
__global__ void __launch_bounds__(32 * (2 + EpiCfg<EPI>::WARPS), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
                    const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const TcArgs args) {
  using SM = GemmSmem<BLOCK_N, EPI, OutT, DUAL, CTAS, NARROW>;
  constexpr int NUM_EPI_WARPS = SM::NUM_EPI_WARPS;
  constexpr int STAGES = SM::STAGES;
  constexpr int BLOCK_K = TILE_K_BYTES / (int)sizeof(InT);     // 64 (bf16) or 32 (tf32)
  constexpr int TMEM_COLS = NUM_ACC_STAGES * BLOCK_N;          // 128 / 256 / 512
  constexpr uint32_t IDESC = umma_idesc(H16<InT>::UMMA_FMT, BLOCK_M * CTAS, BLOCK_N);
  typedef typename Half16Of<InT>::type CopyT;   // type of the 16-bit copy of the residual stream (DUAL)
  // CTA pair: rank within the pair, pair index, number of pairs (CTAS == 1: rank 0, one "pair" per CTA)
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const int pair = (int)blockIdx.x / CTAS;
  const int num_pairs = (int)gridDim.x / CTAS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                            // STAGES x 16 KB
  uint8_t* smem_b = smem + STAGES * A_TILE_BYTES;                    // STAGES x BLOCK_N*128
  uint8_t* staging = smem + STAGES * SM::STAGE_BYTES;                // epilogue staging
  float* bias_smem = reinterpret_cast<float*>(staging + SM::STAGING_BYTES);   // [NUM_EPI_WARPS][BLOCK_N]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + SM::STAGING_BYTES + SM::BIAS_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint64_t* acc_empty = acc_full + NUM_ACC_STAGES;
  uint64_t* res_bar = acc_empty + NUM_ACC_STAGES;                    // [NUM_EPI_WARPS][RES_BUFS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + NUM_EPI_WARPS * RES_BUFS);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  const int num_m = (args.M + BLOCK_M * CTAS - 1) / (BLOCK_M * CTAS);
  const int num_n = args.N / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = args.cin_blocks > 0 ? 9 * args.cin_blocks : (args.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if (DUAL) tma_prefetch_desc(&tmOut2);
    if (EPI == EPI_RESID) tma_prefetch_desc(&tmRes);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < NUM_ACC_STAGES; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], NUM_EPI_WARPS * CTAS);   // the leader's barrier collects both CTAs' epilogue warps
    }
    for (int s = 0; s < NUM_EPI_WARPS * RES_BUFS; ++s) mbar_init(&res_bar[s], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CTAS == 2) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();    // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Optional (GENIE_B200_W_PREFETCH=1, off: measured neutral to slightly negative).  The weights are not produced by
  // the preceding kernel, so their first tiles may be requested while that kernel is still draining (PDL): by the
  // time the dependency resolves, the B operand of every CTA's first tile is in L2.
  if (args.w_prefetch && warp == 0 && lane == 0 && pair < num_tiles) {
    const int n_blk0 = pair % num_n;
    const int kbs = num_kb < 32 ? num_kb : 32;
    for (int kb = 0; kb < kbs; ++kb)
      tma_prefetch_l2_2d(&tmB, kb * BLOCK_K, n_blk0 * BLOCK_N + (int)cta_rank * (BLOCK_N / CTAS));
  }
  pdl_wait();      // everything above overlapped the previous kernel's tail; activations are touched only below
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint64_t pol = l2_policy_evict_first();
      const bool a_hint = args.a_hint != 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        const int m_row0 = (m_blk * CTAS + (int)cta_rank) * BLOCK_M;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (args.dbg & 2) {
            if (cta_rank == 0) mbar_arrive(&full_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if constexpr (CTAS == 2) {
            // both CTAs' bytes complete on the LEADER's barrier (which the MMA issuer waits on)
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * SM::STAGE_BYTES);
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (args.cin_blocks > 0) {
              // conv: this CTA's 128 rows = th x tw output pixels of image n starting at (y0, x0) (same tile geometry
              // as the single-CTA path; the pair shares the weight tile, each CTA staging one half of it)
              const int P = args.Ho * args.Wo;
              const int n = m_row0 / P, rem = m_row0 % P;
              const int y0 = rem / args.Wo, x0 = rem % args.Wo;
              const int tap = kb / args.cin_blocks, cib = kb % args.cin_blocks;
              tma_load_4d_pair(smem_a + stage * A_TILE_BYTES, &tmA, bar, cib * BLOCK_K, x0 * args.stride + tap % 3 - 1,
                               y0 * args.stride + tap / 3 - 1, n);
            } else if (a_hint) tma_load_2d_pair_hint(smem_a + stage * A_TILE_BYTES, &tmA, bar, kb * BLOCK_K, m_row0, pol);
            else tma_load_2d_pair(smem_a + stage * A_TILE_BYTES, &tmA, bar, kb * BLOCK_K, m_row0);
            tma_load_2d_pair(smem_b + stage * SM::B_TILE_BYTES, &tmB, bar, kb * BLOCK_K,
                             n_blk * BLOCK_N + (int)cta_rank * (BLOCK_N / 2));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], SM::STAGE_BYTES);
            if (args.cin_blocks > 0) {
              // conv: tile = th x tw output pixels of image n starting at (y0, x0); k-block = (tap, channel block)
              const int P = args.Ho * args.Wo;
              const int p0 = m_blk * BLOCK_M;
              const int n = p0 / P, rem = p0 % P;
              const int y0 = rem / args.Wo, x0 = rem % args.Wo;
              const int tap = kb / args.cin_blocks, cib = kb % args.cin_blocks;
              tma_load_4d(smem_a + stage * A_TILE_BYTES, &tmA, &full_bar[stage], cib * BLOCK_K,
                          x0 * args.stride + tap % 3 - 1, y0 * args.stride + tap / 3 - 1, n);
            } else if (a_hint) {
              tma_load_2d_hint(smem_a + stage * A_TILE_BYTES, &tmA, &full_bar[stage], kb * BLOCK_K, m_row0, pol);
            } else {
              tma_load_2d(smem_a + stage * A_TILE_BYTES, &tmA, &full_bar[stage], kb * BLOCK_K, m_row0);
            }
            tma_load_2d(smem_b + stage * SM::B_TILE_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair: the leader CTA only)
    if (lane == 0 && cta_rank == 0) {
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(smem_a + stage * A_TILE_BYTES));
          const uint64_t bdesc = umma_desc_kmajor_sw128(smem_u32(smem_b + stage * SM::B_TILE_BYTES));
#pragma unroll
          for (int k = 0; k < TILE_K_BYTES / 32; ++k) {
            // advance 32 B (= one UMMA_K slice) inside the swizzle atom: +2 in 16-byte units
            if constexpr (CTAS == 2) {
              if (sizeof(InT) == 2)
                umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
              else
                umma_tf32_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
            } else {
              if (sizeof(InT) == 2)
                umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
              else
                umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
            }
          }
          // smem slot reusable (in both CTAs) once these MMAs retire
          if constexpr (CTAS == 2) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete (each CTA's epilogue reads its own 128 rows from its own TMEM)
        if constexpr (CTAS == 2) umma_commit_pair(&acc_full[as]);
        else umma_commit(&acc_full[as]);
        if (++as == NUM_ACC_STAGES) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const uint32_t q = warp & 3;                   // TMEM lane quarter this warp may access
    const uint32_t ew = warp - 2;                  // staging slot
    const uint32_t half = ew >> 2;                 // which slice of the tile's columns (8-warp epilogues: 0 / 1)
    constexpr int CH = SM::COLS_PER_WARP / EPI_COLS;   // 32-column chunks per warp per tile
    const int col_base = half * SM::COLS_PER_WARP;
    uint8_t* st0 = staging + ew * SM::OUT_BUFS * SM::OUT_STAGE_BYTES;
    uint8_t* st1 = staging + NUM_EPI_WARPS * SM::OUT_BUFS * SM::OUT_STAGE_BYTES + ew * 2 * SM::OUT2_STAGE_BYTES;
    uint32_t as = 0, aphase = 0;
    float* sbias = bias_smem + ew * SM::COLS_PER_WARP;
    float* scsum = bias_smem + (NUM_EPI_WARPS + ew) * SM::COLS_PER_WARP;
    const bool has_bias = args.bias != nullptr;
    const bool has_ln = args.ln_stats != nullptr;
    // hand an accumulator stage back to the (leader's) MMA issuer
    auto release_acc = [&](uint32_t stage_idx) {
      if (CTAS == 2 && cta_rank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[stage_idx]), 0));
      else mbar_arrive(&acc_empty[stage_idx]);
    };
    float ln_r = 1.f, ln_t = 0.f;   // per-row rstd and -rstd*mean of the folded LayerNorm
    // stage this tile's bias slice in shared memory (per warp) BEFORE waiting for the accumulator, so the
    // global-load latency hides behind the MMA of the tile
    auto stage_bias = [&](int n_blk) {
      if (has_bias || has_ln) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < SM::COLS_PER_WARP / 32; ++i) {
          if (has_bias) sbias[i * 32 + lane] = __ldg(args.bias + n_blk * BLOCK_N + col_base + i * 32 + lane);
          if (has_ln) scsum[i * 32 + lane] = __ldg(args.ln_colsum + n_blk * BLOCK_N + col_base + i * 32 + lane);
        }
        __syncwarp();
      }
    };
    // folded LayerNorm: reduce this row's partial statistics (fixed order) to rstd and -rstd*mean
    auto load_row_stats = [&](int row) {
      if (has_ln) {
        float s1 = 0.f, s2 = 0.f;
        if (row < args.M) {
          const float2* sp = reinterpret_cast<const float2*>(args.ln_stats) + (int64_t)row * args.ln_np;
          for (int i = 0; i < args.ln_np; ++i) { const float2 p = __ldg(sp + i); s1 += p.x; s2 += p.y; }
        }
        const float inv_d = 1.f / (float)args.ln_d;
        const float mean = s1 * inv_d;
        const float var = fmaxf(s2 * inv_d - mean * mean, 0.f);
        ln_r = rsqrtf(var + 1e-5f);
        ln_t = -ln_r * mean;
      }
    };
    auto add_bias = [&](float (&v)[EPI_COLS], int c) {
      if (has_ln) {
        // v = rstd * acc + (-rstd * mean) * colsum[n]   (then the bias, which already holds beta . W^T)
        const float4* cp = reinterpret_cast<const float4*>(scsum + c * EPI_COLS);
#pragma unroll
        for (int j = 0; j < EPI_COLS / 4; ++j) {
          const float4 cs = cp[j];
          v[4 * j] = fmaf(v[4 * j], ln_r, ln_t * cs.x);
          v[4 * j + 1] = fmaf(v[4 * j + 1], ln_r, ln_t * cs.y);
          v[4 * j + 2] = fmaf(v[4 * j + 2], ln_r, ln_t * cs.z);
          v[4 * j + 3] = fmaf(v[4 * j + 3], ln_r, ln_t * cs.w);
        }
      }
      if (has_bias) {
        const float4* bp = reinterpret_cast<const float4*>(sbias + c * EPI_COLS);
#pragma unroll
        for (int j = 0; j < EPI_COLS / 4; ++j) {
          const float4 b = bp[j];
          v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
      }
    };
    if constexpr (EPI == EPI_RESID) {
      // Residual epilogue.  The fp32 residual chunk (32 rows x 32 cols of this warp) is TMA-loaded into a
      // swizzled staging buffer RES_PREFETCH chunks ahead, updated IN PLACE with acc + bias, and TMA-stored
      // from the same buffer: every global access of the epilogue is a coalesced bulk copy.
      constexpr int RB = SM::OUT_BUFS;   // staging buffers per warp
      constexpr int RP = RB - 2;         // residual chunks requested ahead of use
      uint64_t* rbar = res_bar + ew * RES_BUFS;
      const int my_tiles = pair < num_tiles ? (num_tiles - pair + num_pairs - 1) / num_pairs : 0;
      const int total_chunks = my_tiles * CH;
      auto issue_residual = [&](int g) {           // lane 0 only
        const int t = pair + (g / CH) * num_pairs;
        const int mb = t / num_n, nb = t % num_n;
        const int b = g % RB;
        mbar_arrive_expect_tx(&rbar[b], SM::OUT_STAGE_BYTES);
        tma_load_2d(st0 + b * SM::OUT_STAGE_BYTES, &tmRes, &rbar[b], nb * BLOCK_N + (g % CH) * EPI_COLS,
                    (mb * CTAS + (int)cta_rank) * BLOCK_M + q * 32);
      };
      if (lane == 0) {
        for (int g0 = 0; g0 < RP && g0 < total_chunks; ++g0) issue_residual(g0);
      }
      int g = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        const int row0 = (m_blk * CTAS + (int)cta_rank) * BLOCK_M + q * 32;
        stage_bias(n_blk);
        mbar_wait(&acc_full[as], aphase);
        tc_fence_after();
        float rs1 = 0.f, rs2 = 0.f;   // partial {sum, sumsq} of this thread's row over the tile's columns
        const uint32_t acc_addr = tmem_base + ((q * 32u) << 16) + as * BLOCK_N;
        uint32_t rr[2][EPI_COLS];
        tmem_ld_32x32b_x32(acc_addr, rr[0]);
        // one chunk: wait for its TMEM load and start the next chunk's load (it overlaps the residual read-modify-write
        // below), then acc + bias + residual in place in the staging buffer, TMA store
        auto do_chunk = [&](uint32_t (&r)[EPI_COLS], uint32_t (&rnext)[EPI_COLS], int c) {
          const int col0 = n_blk * BLOCK_N + c * EPI_COLS;
          if (lane == 0) {
            // buffer (g + RP) % RB was last stored from RB - RP chunks ago
            tma_store_wait_read<RB - RP - 1>();
            if (g + RP < total_chunks) issue_residual(g + RP);
          }
          tmem_ld_wait();
          if (c + 1 < CH) tmem_ld_32x32b_x32(acc_addr + (c + 1) * EPI_COLS, rnext);
          if (c == CH - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(as);
          }
          float v[EPI_COLS];
#pragma unroll
          for (int j = 0; j < EPI_COLS; ++j) v[j] = __uint_as_float(r[j]);
          add_bias(v, c);
          const int b = g % RB;
          mbar_wait(&rbar[b], (g / RB) & 1);
          __syncwarp();                             // lane 0's wait_read above precedes every lane's staging writes
          uint8_t* rowp = st0 + b * SM::OUT_STAGE_BYTES + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4* p = reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4));
            const float4 x4 = *p;
            v[4 * j] += x4.x; v[4 * j + 1] += x4.y; v[4 * j + 2] += x4.z; v[4 * j + 3] += x4.w;
            *p = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (args.stats_out != nullptr) {
#pragma unroll
            for (int j = 0; j < EPI_COLS; ++j) { rs1 += v[j]; rs2 = fmaf(v[j], v[j], rs2); }
          }
          if (DUAL) stage_row_chunk<CopyT>(st1 + (g & 1) * SM::OUT2_STAGE_BYTES, lane, v);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmOut, st0 + b * SM::OUT_STAGE_BYTES, col0, row0);
            if (DUAL) tma_store_2d(&tmOut2, st1 + (g & 1) * SM::OUT2_STAGE_BYTES, col0, row0);
            tma_store_commit();
          }
          ++g;
        };
#pragma unroll 1
        for (int c2 = 0; c2 < CH; c2 += 2) {
          do_chunk(rr[0], rr[1], c2);
          if (c2 + 1 < CH) do_chunk(rr[1], rr[0], c2 + 1);
        }
        if (args.stats_out != nullptr && row0 + (int)lane < args.M) {
          reinterpret_cast<float2*>(args.stats_out)[(int64_t)(row0 + lane) * num_n + n_blk] = make_float2(rs1, rs2);
        }
        if (++as == NUM_ACC_STAGES) { as = 0; aphase ^= 1; }
      }
    } else {
      uint32_t buf = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        const int row0 = (m_blk * CTAS + (int)cta_rank) * BLOCK_M + q * 32;
        stage_bias(n_blk);
        load_row_stats(row0 + (int)lane);
        mbar_wait(&acc_full[as], aphase);
        tc_fence_after();
        if (args.dbg & 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(as);
          if (++as == NUM_ACC_STAGES) { as = 0; aphase ^= 1; }
          continue;
        }
        const uint32_t acc_addr = tmem_base + ((q * 32u) << 16) + as * BLOCK_N + col_base;
        // K/V-cache destination of this warp's 32 rows (rows are (clip, local frame, s); 32 | S)
        int kv_frame = 0, kv_pos = 0;
        if (EPI == EPI_STORE && sizeof(OutT) == 2 && args.kv_d > 0) {
          const int bt = row0 / args.kv_S;
          kv_frame = args.kv_t0 + bt % args.kv_Tact;
          kv_pos = (bt / args.kv_Tact) * args.kv_S + row0 % args.kv_S;
        }
        uint32_t rr[2][EPI_COLS];
        float vsave[QKN ? EPI_COLS : 1];   // first half of a 64-column pair, kept for the qk-LayerNorm of the pair
        tmem_ld_32x32b_x32(acc_addr, rr[0]);
        // one chunk: wait for its TMEM load, start the next chunk's load (overlaps the math), bias / activation,
        // swizzled staging, TMA store
        auto do_chunk = [&](uint32_t (&r)[EPI_COLS], uint32_t (&rnext)[EPI_COLS], int c) {
          const int col0 = n_blk * BLOCK_N + col_base + c * EPI_COLS;
          tmem_ld_wait();
          if (c + 1 < CH) tmem_ld_32x32b_x32(acc_addr + (c + 1) * EPI_COLS, rnext);
          if (c == CH - 1) {
            // accumulator stage fully read into registers: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(as);
          }
          float v[EPI_COLS];
#pragma unroll
          for (int j = 0; j < EPI_COLS; ++j) v[j] = __uint_as_float(r[j]);
          add_bias(v, c);
          if (EPI == EPI_GELU) {
            if (sizeof(OutT) == 2) {
#pragma unroll
              for (int j = 0; j < EPI_COLS; j += 2) gelu_fast2(v[j], v[j + 1]);
            } else {
#pragma unroll
              for (int j = 0; j < EPI_COLS; ++j) v[j] = gelu_erf(v[j]);
            }
          }
          if (sizeof(OutT) == 4 && args.round_tf32) {
#pragma unroll
            for (int j = 0; j < EPI_COLS; ++j) v[j] = tf32_rn(v[j]);
          }
          if constexpr (SM::WIDE) {
            const int half = c & 1;
            // the staging buffer about to be written was last stored from OUT_BUFS stores ago
            if (half == 0) {
              if (lane == 0) tma_store_wait_read<SM::OUT_BUFS - 1>();
              __syncwarp();
            }
            // qk-LayerNorm: the two chunks of a pair are the 64 columns of one head of this thread's row
            const bool qkn = QKN && (col0 - half * EPI_COLS) < args.qkn_cols;
            if constexpr (!QKN) {
              if (!(args.dbg & 8)) stage_row_chunk_wide<OutT>(st0 + buf * SM::OUT_STAGE_BYTES, lane, half, v);
            } else if (qkn) {
              if (half == 0) {
#pragma unroll
                for (int j = 0; j < EPI_COLS; ++j) vsave[j] = v[j];
              } else {
                float s1 = 0.f;
#pragma unroll
                for (int j = 0; j < EPI_COLS; ++j) s1 += vsave[j] + v[j];
                const float mean = s1 * (1.f / (2 * EPI_COLS));
                float s2 = 0.f;
#pragma unroll
                for (int j = 0; j < EPI_COLS; ++j) {
                  const float d0 = vsave[j] - mean, d1 = v[j] - mean;
                  s2 = fmaf(d0, d0, s2);
                  s2 = fmaf(d1, d1, s2);
                }
                const float rstd = rsqrtf(s2 * (1.f / (2 * EPI_COLS)) + 1e-5f);
                const float4* g4 = reinterpret_cast<const float4*>(args.qkn_g);
                const float4* b4 = reinterpret_cast<const float4*>(args.qkn_b);
#pragma unroll
                for (int j = 0; j < EPI_COLS / 4; ++j) {
                  const float4 ga = __ldg(g4 + j), ba = __ldg(b4 + j);
                  const float4 gb = __ldg(g4 + EPI_COLS / 4 + j), bb = __ldg(b4 + EPI_COLS / 4 + j);
                  vsave[4 * j] = (vsave[4 * j] - mean) * rstd * ga.x + ba.x;
                  vsave[4 * j + 1] = (vsave[4 * j + 1] - mean) * rstd * ga.y + ba.y;
                  vsave[4 * j + 2] = (vsave[4 * j + 2] - mean) * rstd * ga.z + ba.z;
                  vsave[4 * j + 3] = (vsave[4 * j + 3] - mean) * rstd * ga.w + ba.w;
                  v[4 * j] = (v[4 * j] - mean) * rstd * gb.x + bb.x;
                  v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * gb.y + bb.y;
                  v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * gb.z + bb.z;
                  v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * gb.w + bb.w;
                }
                stage_row_chunk_wide<OutT>(st0 + buf * SM::OUT_STAGE_BYTES, lane, 0, vsave);
                stage_row_chunk_wide<OutT>(st0 + buf * SM::OUT_STAGE_BYTES, lane, 1, v);
              }
            } else if (!(args.dbg & 8)) {
              stage_row_chunk_wide<OutT>(st0 + buf * SM::OUT_STAGE_BYTES, lane, half, v);
            }
            if (half == 1) {
              fence_proxy_async_smem();
              __syncwarp();
              const int colp = col0 - EPI_COLS;   // first column of the pair
              if (lane == 0 && !(args.dbg & 4)) {
                if (EPI == EPI_STORE && args.kv_d > 0 && colp >= args.kv_d) {
                  const int part = colp >= 2 * args.kv_d ? 2 : 1;
                  const int cc = colp - part * args.kv_d;
                  tma_store_4d(part == 1 ? &tmK : &tmV, st0 + buf * SM::OUT_STAGE_BYTES, cc % args.kv_hd, kv_frame,
                               cc / args.kv_hd, kv_pos);
                } else {
                  tma_store_2d(&tmOut, st0 + buf * SM::OUT_STAGE_BYTES, colp, row0);
                }
                tma_store_commit();
              }
              buf = (buf + 1 == SM::OUT_BUFS) ? 0 : buf + 1;
            }
          } else {
            if constexpr (QKN && NARROW) {
              // qk-LayerNorm for head_dim 32: this 32-column chunk is one head of this thread's row
              if (col0 < args.qkn_cols) {
                float s1 = 0.f;
#pragma unroll
                for (int j = 0; j < EPI_COLS; ++j) s1 += v[j];
                const float mean = s1 * (1.f / EPI_COLS);
                float s2 = 0.f;
#pragma unroll
                for (int j = 0; j < EPI_COLS; ++j) { const float dl = v[j] - mean; s2 = fmaf(dl, dl, s2); }
                const float rstd = rsqrtf(s2 * (1.f / EPI_COLS) + 1e-5f);
                const float4* g4 = reinterpret_cast<const float4*>(args.qkn_g);
                const float4* b4 = reinterpret_cast<const float4*>(args.qkn_b);
#pragma unroll
                for (int j = 0; j < EPI_COLS / 4; ++j) {
                  const float4 ga = __ldg(g4 + j), ba = __ldg(b4 + j);
                  v[4 * j] = (v[4 * j] - mean) * rstd * ga.x + ba.x;
                  v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * ga.y + ba.y;
                  v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * ga.z + ba.z;
                  v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * ga.w + ba.w;
                }
              }
            }
            // the staging buffer about to be written was last stored from OUT_BUFS steps ago
            if (lane == 0) tma_store_wait_read<SM::OUT_BUFS - 1>();
            __syncwarp();
            if (!(args.dbg & 8)) stage_row_chunk<OutT>(st0 + buf * SM::OUT_STAGE_BYTES, lane, v);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(args.dbg & 4)) {
              if (EPI == EPI_STORE && sizeof(OutT) == 2 && args.kv_d > 0 && col0 >= args.kv_d) {
                const int part = col0 >= 2 * args.kv_d ? 2 : 1;
                const int cc = col0 - part * args.kv_d;
                tma_store_4d(part == 1 ? &tmK : &tmV, st0 + buf * SM::OUT_STAGE_BYTES, cc % args.kv_hd, kv_frame,
                             cc / args.kv_hd, kv_pos);
              } else if (sizeof(OutT) == 4 && args.red_add) {
                tma_reduce_add_2d(&tmOut, st0 + buf * SM::OUT_STAGE_BYTES, col0, row0);
              } else {
                tma_store_2d(&tmOut, st0 + buf * SM::OUT_STAGE_BYTES, col0, row0);
              }
              tma_store_commit();
            }
            if (SM::OUT_BUFS > 1) buf = (buf + 1 == SM::OUT_BUFS) ? 0 : buf + 1;
          }
        };
#pragma unroll 1
        for (int c2 = 0; c2 < CH; c2 += 2) {
          do_chunk(rr[0], rr[1], c2);
          if (c2 + 1 < CH) do_chunk(rr[1], rr[0], c2 + 1);
        }
        if (++as == NUM_ACC_STAGES) { as = 0; aphase ^= 1; }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();    // neither CTA may retire while the pair's MMAs / commits still target it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <typename InT, int BLOCK_N, int EPI, typename OutT, bool DUAL, int CTAS = 1, bool QKN = false,
          bool NARROW = false>
int launch_tc(const LinearArgs& a, cudaStream_t stream) {
  using SM = GemmSmem<BLOCK_N, EPI, OutT, DUAL, CTAS, NARROW>;
  constexpr int BLOCK_K = TILE_K_BYTES / (int)sizeof(InT);
  const CUtensorMapDataType in_dt = H16<InT>::TMAP;
  const CUtensorMapDataType out_dt = H16<OutT>::TMAP;
  const CUtensorMapDataType copy_dt = H16<typename Half16Of<InT>::type>::TMAP;
  CUtensorMap tmA, tmB, tmO, tmO2, tmR, tmK, tmV;
  int tw = 0, th = 0;
  if (a.conv) {
    const ConvGeom& g = *a.conv;
    tw = g.Wo < BLOCK_M ? g.Wo : BLOCK_M;
    th = BLOCK_M / tw;
    GN_PROPAGATE(make_tensor_map_nhwc(&tmA, a.A, in_dt, sizeof(InT), g.Nimg, g.Hi, g.Wi, g.Cin, BLOCK_K, tw, th, g.stride,
                                      g.stride, CU_TENSOR_MAP_SWIZZLE_128B));
  } else
  GN_PROPAGATE(make_tensor_map_2d(&tmA, a.A, in_dt, sizeof(InT), a.K, a.M, a.lda, BLOCK_K, BLOCK_M,
                                  CU_TENSOR_MAP_SWIZZLE_128B));
  GN_PROPAGATE(make_tensor_map_2d(&tmB, a.W, in_dt, sizeof(InT), a.K, a.N, a.ldw, BLOCK_K, BLOCK_N / CTAS,
                                  CU_TENSOR_MAP_SWIZZLE_128B));
  GN_PROPAGATE(make_tensor_map_2d(&tmO, a.out, out_dt, sizeof(OutT), a.N, a.M, a.ldo,
                                  SM::WIDE ? 2 * EPI_COLS : EPI_COLS, 32,
                                  (sizeof(OutT) == 2 && !SM::WIDE) ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                   : CU_TENSOR_MAP_SWIZZLE_128B));
  if (DUAL) {
    GN_PROPAGATE(make_tensor_map_2d(&tmO2, a.out2, copy_dt, 2, a.N, a.M, a.ldo2, EPI_COLS, 32,
                                    CU_TENSOR_MAP_SWIZZLE_64B));
  } else {
    tmO2 = tmO;
  }
  if (EPI == EPI_RESID) {
    GN_PROPAGATE(make_tensor_map_2d(&tmR, a.resid, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.N, a.M, a.ldr, EPI_COLS, 32,
                                    CU_TENSOR_MAP_SWIZZLE_128B));
  } else {
    tmR = tmO;
  }
  const bool kv = EPI == EPI_STORE && sizeof(OutT) == 2 && a.kv_k != nullptr;
  if (kv) {
    // cache view [clip*S + s][head][T][hd]; one store = 32 positions x 32 columns of one head and frame
    const int H = a.kv_d / a.kv_hd;
    const int64_t dims[4] = {a.kv_hd, a.kv_T, H, (int64_t)a.kv_clips * a.kv_S};
    const int64_t str[3] = {(int64_t)a.kv_hd * 2, (int64_t)a.kv_T * a.kv_hd * 2, (int64_t)H * a.kv_T * a.kv_hd * 2};
    // wide staging: one store = 32 positions x one full head (64 columns = 128 B) of one frame
    if (SM::WIDE && a.kv_hd != 2 * EPI_COLS) {
      set_error("K/V-cache epilogue with wide staging needs head_dim %d (got %d)", 2 * EPI_COLS, a.kv_hd);
      return GN_ERR_UNSUPPORTED;
    }
    const int box[4] = {SM::WIDE ? 2 * EPI_COLS : EPI_COLS, 1, 1, 32};
    const CUtensorMapSwizzle ksw = SM::WIDE ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    GN_PROPAGATE(make_tensor_map_nd(&tmK, a.kv_k, out_dt, 4, dims, str, box, ksw));
    GN_PROPAGATE(make_tensor_map_nd(&tmV, a.kv_v, out_dt, 4, dims, str, box, ksw));
  } else {
    tmK = tmO;
    tmV = tmO;
  }
  auto kern = gemm_tcgen05_kernel<InT, BLOCK_N, EPI, OutT, DUAL, CTAS, QKN, NARROW>;
  static DevSmemOptIn optin;
  GN_CUDA_CHECK(ensure_smem_optin(optin, kern, SM::TOTAL));
  const int num_tiles = ceil_div(a.M, BLOCK_M * CTAS) * (a.N / BLOCK_N);
  const int sms = device_sm_count();
  int grid = num_tiles < sms ? num_tiles : sms;
  if (CTAS == 2) {
    // persistent pairs: as many 2-CTA clusters as can be co-resident (one CTA per SM, both SMs of a TPC)
    static int max_pairs_dev[kMaxDevices] = {};
    int& max_pairs = max_pairs_dev[current_device()];
    if (!max_pairs) {
      cudaLaunchConfig_t qc{};
      qc.gridDim = dim3(sms);
      qc.blockDim = dim3(32 * (2 + SM::NUM_EPI_WARPS));
      qc.dynamicSmemBytes = SM::TOTAL;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      qc.attrs = qa;
      qc.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &qc) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = sms / 2;
      }
      max_pairs = n < sms / 2 ? n : sms / 2;
    }
    grid = 2 * (num_tiles < max_pairs ? num_tiles : max_pairs);
  }
  TcArgs t{a.M, a.N, a.K, a.bias, a.resid, a.ldr, a.round_out_tf32, 0, 0, 0, 0, 0, 1, nullptr, 0, 0, nullptr, nullptr,
           0, 0, 0, 0, 0, 0};
  {
    const char* e = getenv("GENIE_B200_GEMM_DEBUG");   // re-read per launch: scripts/gemm_ablation.py flips it
    t.dbg = e ? atoi(e) : 0;
  }
  t.a_hint = a.a_evict_first;
  t.qkn_g = a.qkn_gamma; t.qkn_b = a.qkn_beta; t.qkn_cols = a.qkn_gamma ? a.qkn_cols : 0;
  t.red_add = a.red_add;
  if ((t.qkn_cols > 0) != QKN || (QKN && !((SM::WIDE || NARROW) && EPI == EPI_STORE && sizeof(OutT) == 2))) {
    set_error("qk-LayerNorm epilogue needs the 16-bit store epilogue (head_dim 64: wide staging, N %% 128 == 0; "
              "head_dim 32: narrow staging)");
    return GN_ERR_UNSUPPORTED;
  }
  t.w_prefetch = env_on("GENIE_B200_W_PREFETCH", false) ? 1 : 0;   // measured: 257.9 vs 257.3 ms per step -> off
  if (kv) { t.kv_d = a.kv_d; t.kv_hd = a.kv_hd; t.kv_S = a.kv_S; t.kv_Tact = a.kv_Tact; t.kv_t0 = a.kv_t0; }
  if (a.conv) {
    t.cin_blocks = a.conv->Cin / BLOCK_K;
    t.Ho = a.conv->Ho; t.Wo = a.conv->Wo; t.tw = tw; t.th = th; t.stride = a.conv->stride;
  }
  t.ln_stats = a.ln_stats; t.ln_np = a.ln_np; t.ln_d = a.ln_d; t.ln_colsum = a.ln_colsum; t.stats_out = a.stats_out;
  g_gemm_flops_issued += 2.0 * a.M * (double)a.N * a.K;
  const int cat = (EPI == EPI_RESID || a.red_add) ? PC_GEMM_RESID : (EPI == EPI_GELU ? PC_GEMM_GELU : PC_GEMM_STORE);
  if (CTAS == 2)
    GN_CUDA_CHECK(launch_kernel_cluster(cat, kern, dim3(grid), dim3(32 * (2 + SM::NUM_EPI_WARPS)), SM::TOTAL, stream, 2,
                                        tmA, tmB, tmO, tmO2, tmR, tmK, tmV, t));
  else
    GN_CUDA_CHECK(launch_kernel(cat, kern, dim3(grid), dim3(32 * (2 + SM::NUM_EPI_WARPS)), SM::TOTAL, stream, tmA, tmB,
                                tmO, tmO2, tmR, tmK, tmV, t));
  ++g_launch_count;
  return GN_OK;
}

template <typename InT, int BLOCK_N, int CTAS = 1>
int dispatch_epi(const LinearArgs& a, cudaStream_t s) {
  typedef typename Half16Of<InT>::type O16;   // 16-bit outputs share the operands' 16-bit format
  if (a.epi == EPI_STORE) {
    if constexpr (sizeof(InT) == 2) {   // head_dim 32: 32-column (64-byte) staging / K/V-cache lines, per-chunk qk-LN
      if (a.out_bf16 && a.qkn_gamma && a.qkn_hd == EPI_COLS)
        return launch_tc<InT, BLOCK_N, EPI_STORE, O16, false, CTAS, true, true>(a, s);
      if (a.out_bf16 && a.kv_k && a.kv_hd == EPI_COLS)
        return launch_tc<InT, BLOCK_N, EPI_STORE, O16, false, CTAS, false, true>(a, s);
    }
    if constexpr (sizeof(InT) == 2 && BLOCK_N >= 128) {
      if (a.out_bf16 && a.qkn_gamma) return launch_tc<InT, BLOCK_N, EPI_STORE, O16, false, CTAS, true>(a, s);
    }
    return a.out_bf16 ? launch_tc<InT, BLOCK_N, EPI_STORE, O16, false, CTAS>(a, s)
                      : launch_tc<InT, BLOCK_N, EPI_STORE, float, false, CTAS>(a, s);
  }
  if (a.epi == EPI_GELU) {
    return a.out_bf16 ? launch_tc<InT, BLOCK_N, EPI_GELU, O16, false, CTAS>(a, s)
                      : launch_tc<InT, BLOCK_N, EPI_GELU, float, false, CTAS>(a, s);
  }
  if (a.epi == EPI_RESID) {
    if (a.out_bf16) { set_error("EPI_RESID writes the fp32 residual stream"); return GN_ERR_INVALID; }
    if constexpr (BLOCK_N <= 128) {
      if constexpr (CTAS == 2 && BLOCK_N == 128) {   // CTA-pair convolution tiles (256 pixels x 128 channels)
        if (a.conv && !a.out2) return launch_tc<InT, BLOCK_N, EPI_RESID, float, false, 2>(a, s);
      }
      return a.out2 ? launch_tc<InT, BLOCK_N, EPI_RESID, float, true>(a, s)
                    : launch_tc<InT, BLOCK_N, EPI_RESID, float, false>(a, s);
    } else {   // 256-wide: 3 residual staging buffers when the bf16 copy is emitted too
      return a.out2 ? launch_tc<InT, BLOCK_N, EPI_RESID, float, true, CTAS>(a, s)
                    : launch_tc<InT, BLOCK_N, EPI_RESID, float, false, CTAS>(a, s);
    }
  }
  set_error("unknown epilogue %d", a.epi);
  return GN_ERR_INVALID;
}

template <typename InT>
int dispatch_n(const LinearArgs& a, cudaStream_t s) {
  // BLOCK_N = 256 keeps the smem operand traffic per MMA cycle under the 128 B/clk port limit.
  // Residual epilogues stage 4 fp32 buffers per warp, so they normally take BLOCK_N = 128 to keep a deep TMA ring;
  // for long-K residual GEMMs (fc2: K = 4d) the 128-wide tile is operand-bandwidth bound (A+B = 128 B/clk of smem
  // reads per MMA cycle), so those use 256 with a 3-stage ring.
  // 256-wide tiles run as CTA pairs (256 x 256 per pair, cta_group::2) unless disabled or a convolution
  const bool pair = env_on("GENIE_B200_PAIR", g_use_pair) && !a.conv && a.M > BLOCK_M;
  // Convolutions (MAGVIT2): CTA-pair tiles of 256 pixels x BLOCK_N channels, each CTA staging its own 128-pixel A box and
  // half of the weight tile.  The 3x3 convs at Cout = 128 move 16 KB (A) + 16 KB (B) from L2 per 2 MFLOP k-block in the
  // single-CTA form - twice the L2 -> SM intensity of the K = 512 linear layers, which is what bounds them; the pair
  // form needs 16 + 8 KB.  MEASURED AND LEFT OFF (GENIE_B200_CONV_PAIR=1 enables it; tokenizer tests pass with it): the
  // Cout = 128 convs got slower (ncu: 464 -> 528 us store, 553 -> 602 us residual per launch at 32 images), the Cout >= 256
  // ones 114 -> 100 us; encode 3187-3241 vs 3161-3171 img/s, decode 2328-2341 vs 2330-2332: the 128-wide conv tiles are
  // not bound by the L2 -> SM operand stream.
  if (a.conv && a.M % (2 * BLOCK_M) == 0 && env_on("GENIE_B200_PAIR", g_use_pair) && env_on("GENIE_B200_CONV_PAIR", false)) {
    if (a.N % 256 == 0) return dispatch_epi<InT, 256, 2>(a, s);
    if (a.N % 128 == 0) return dispatch_epi<InT, 128, 2>(a, s);
  }
  // experiment switch for the N = 512 GEMMs (proj / fc2): GENIE_B200_BN512 = 64 | 128 | 256 (single CTA) | 1128 | 1256
  // (CTA pair with 128 / 256-wide tiles)
  if (a.N == 512 && !a.conv) {
    const char* e = getenv("GENIE_B200_BN512");
    const int v = e ? atoi(e) : 0;
    if (v == 64) return dispatch_epi<InT, 64>(a, s);
    if (v == 128) return dispatch_epi<InT, 128>(a, s);
    if (v == 256) return dispatch_epi<InT, 256>(a, s);
    if (v == 1128 && a.M > BLOCK_M) return dispatch_epi<InT, 128, 2>(a, s);
    if (v == 1256 && a.M > BLOCK_M) return dispatch_epi<InT, 256, 2>(a, s);
  }
  if (a.N % 256 == 0 && pair && (a.epi != EPI_RESID || a.K >= 1024 || env_on("GENIE_B200_PAIR_PROJ", false)))
    return dispatch_epi<InT, 256, 2>(a, s);
  if (a.N % 256 == 0 && (a.epi != EPI_RESID || a.K >= 1024 || env_on("GENIE_B200_PROJ256", false)))
    return dispatch_epi<InT, 256>(a, s);
  if (a.N % 128 == 0) return dispatch_epi<InT, 128>(a, s);
  return dispatch_epi<InT, 64>(a, s);
}

// =====================================================================================
// CUDA-core fp32 path (generic shapes, "exact" mode)
// =====================================================================================
constexpr int ST = 64, SK = 16;

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const InT* __restrict__ A, int64_t lda, const InT* __restrict__ W, int64_t ldw,
                 const float* __restrict__ bias, const float* __restrict__ resid, int64_t ldr, OutT* __restrict__ out,
                 int64_t ldo, typename Half16Of<InT>::type* __restrict__ out2, int64_t ldo2, int M, int N, int K,
                 int epi, ConvGeom cg) {
  __shared__ float sa[SK][ST + 1];
  __shared__ float sw[SK][ST + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * ST, n0 = blockIdx.x * ST;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SK) {
    for (int i = threadIdx.x; i < ST * SK; i += 256) {
      const int r = i / SK, c = i % SK;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + c;
      float av = 0.f;
      if (gm < M && gk < K) {
        if (cg.Cin > 0) {
          // implicit GEMM 3x3 convolution, padding 1: row = output pixel (n, y, x), column = (tap, channel)
          const int P = cg.Ho * cg.Wo;
          const int n = gm / P, rem = gm % P, y = rem / cg.Wo, x = rem % cg.Wo;
          const int tap = gk / cg.Cin, ci = gk % cg.Cin;
          const int iy = y * cg.stride + tap / 3 - 1, ix = x * cg.stride + tap % 3 - 1;
          if (iy >= 0 && iy < cg.Hi && ix >= 0 && ix < cg.Wi)
            av = to_f32<InT>(A[(((int64_t)n * cg.Hi + iy) * cg.Wi + ix) * cg.Cin + ci]);
        } else {
          av = to_f32<InT>(A[(int64_t)gm * lda + gk]);
        }
      }
      sa[c][r] = av;
      sw[c][r] = (gn < N && gk < K) ? to_f32<InT>(W[(int64_t)gn * ldw + gk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sa[k][ty * 4 + i]; wv[i] = sw[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      if (epi == EPI_GELU) v = gelu_erf(v);
      if (epi == EPI_RESID) v += resid[(int64_t)gm * ldr + gn];
      out[(int64_t)gm * ldo + gn] = from_f32<OutT>(v);
      if (out2) out2[(int64_t)gm * ldo2 + gn] = from_f32<typename Half16Of<InT>::type>(v);
    }
  }
}

template <typename InT, typename OutT>
int launch_simt(const LinearArgs& a, cudaStream_t s) {
  dim3 grid(ceil_div(a.N, ST), ceil_div(a.M, ST));
  gemm_simt_kernel<InT, OutT><<<grid, 256, 0, s>>>(
      static_cast<const InT*>(a.A), a.lda, static_cast<const InT*>(a.W), a.ldw, a.bias, a.resid, a.ldr,
      static_cast<OutT*>(a.out), a.ldo, static_cast<typename Half16Of<InT>::type*>(a.out2), a.ldo2, a.M, a.N, a.K,
      a.epi, a.conv ? *a.conv : ConvGeom{0, 0, 0, 0, 0, 0, 1});
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

}  // namespace

int resid_block_n(int N, int K, bool dual) {
  if (N % 256 == 0 && env_on("GENIE_B200_PAIR", g_use_pair) && env_on("GENIE_B200_PAIR_PROJ", false)) return 256;
  if (N % 256 == 0 && K >= 1024) return 256;
  if (N % 128 == 0) return 128;
  return 64;
}

int linear_forward(const LinearArgs& a, cudaStream_t stream) {
  GN_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "linear_forward: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  GN_REQUIRE(a.A && a.W && a.out, "linear_forward: null operand");
  GN_REQUIRE(a.epi != EPI_RESID || a.resid, "linear_forward: EPI_RESID needs a residual pointer");
  GN_REQUIRE(!a.red_add || (a.epi == EPI_STORE && !a.out_bf16 && !a.force_simt && !a.conv && !a.round_out_tf32),
             "red_add: fp32 store epilogue on the tensor path only");
  const int esz = a.in_bf16 ? 2 : 4;
  GN_REQUIRE(!(a.ln_stats || a.stats_out) || (!a.force_simt && a.N % 64 == 0),
             "folded LayerNorm / row statistics need the tensor path (N %% 64 == 0)");
  GN_REQUIRE(!a.ln_stats || (a.ln_colsum && a.ln_np > 0 && a.ln_d > 0 && a.epi != EPI_RESID), "bad folded-LayerNorm arguments");
  GN_REQUIRE(!a.stats_out || a.epi == EPI_RESID, "row statistics are produced by the residual epilogue");
  if (a.conv) {
    const ConvGeom& g = *a.conv;
    GN_REQUIRE(a.K == 9 * g.Cin, "conv: K == 9*Cin expected");
    GN_REQUIRE(g.stride == 1 || g.stride == 2, "conv: stride must be 1 or 2");
    GN_REQUIRE(a.M == g.Nimg * g.Ho * g.Wo, "conv: M != Nimg*Ho*Wo");
  }
  if (a.conv && !a.force_simt) {   // the fp32 exact mode (force_simt) takes any geometry on the CUDA-core kernel
    const ConvGeom& g = *a.conv;
    const int bk = 128 / esz;
    GN_REQUIRE(g.Cin % bk == 0, "conv: Cin %d must be a multiple of %d", g.Cin, bk);
    GN_REQUIRE(a.N % 64 == 0, "conv: Cout %d must be a multiple of 64", a.N);
    GN_REQUIRE((g.Ho * g.Wo) % 128 == 0 && (g.Wo >= 128 ? g.Wo % 128 == 0 : 128 % g.Wo == 0),
               "conv: output %dx%d not tileable by 128-pixel tiles", g.Ho, g.Wo);
    GN_REQUIRE(g.Wo >= 128 || g.Ho % (128 / g.Wo) == 0, "conv: Ho not a multiple of the tile height");
    GN_REQUIRE((g.Wo < 128 ? g.Wo : 128) * g.stride <= 256, "conv: TMA box too wide");
  }
  if (a.kv_k) {
    GN_REQUIRE(a.kv_v && a.epi == EPI_STORE && a.in_bf16 && a.out_bf16 && !a.force_simt && !a.conv,
               "K/V-cache output needs the bf16 tensor path with the store epilogue");
    GN_REQUIRE(a.kv_d > 0 && a.N == 3 * a.kv_d && a.kv_hd % EPI_COLS == 0 && a.kv_d % a.kv_hd == 0 && a.kv_S % 32 == 0 &&
                   a.kv_Tact > 0 && a.kv_t0 >= 0 && a.kv_t0 + a.kv_Tact <= a.kv_T && a.M == a.kv_clips * a.kv_Tact * a.kv_S,
               "K/V-cache output: inconsistent geometry");
  }
  if (a.qkn_gamma) {
    GN_REQUIRE(a.qkn_beta && a.epi == EPI_STORE && a.in_bf16 && a.out_bf16 && !a.force_simt && !a.conv &&
                   (a.qkn_hd == 2 * EPI_COLS || a.qkn_hd == EPI_COLS) && a.qkn_cols > 0 && a.qkn_cols % a.qkn_hd == 0 &&
                   a.qkn_cols <= a.N && (a.qkn_hd == EPI_COLS || a.N % 128 == 0) && !a.ln_stats,
               "qk-LayerNorm epilogue: 16-bit store epilogue on the tensor path, head_dim 64 (N %% 128 == 0) or 32");
  }
  const bool tc_ok = !a.force_simt && (a.N % 64 == 0) && (a.K * esz % 16 == 0) && (a.lda * esz % 16 == 0) &&
                     (a.ldw * esz % 16 == 0) && (a.ldo * (a.out_bf16 ? 2 : 4) % 16 == 0) &&
                     (!a.out2 || a.ldo2 * 2 % 16 == 0) && (!a.resid || a.ldr * 4 % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(a.A) % 16 == 0) && (reinterpret_cast<uintptr_t>(a.W) % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(a.out) % 16 == 0) &&
                     (!a.out2 || reinterpret_cast<uintptr_t>(a.out2) % 16 == 0) &&
                     (!a.resid || reinterpret_cast<uintptr_t>(a.resid) % 16 == 0) &&
                     (!a.bias || reinterpret_cast<uintptr_t>(a.bias) % 16 == 0) &&
                     !(a.out2 && (a.epi != EPI_RESID || a.out_bf16));
  if (tc_ok) {
    if (!a.in_bf16) return dispatch_n<float>(a, stream);
    return a.fp16 ? dispatch_n<f16>(a, stream) : dispatch_n<bf16>(a, stream);
  }
  GN_REQUIRE(!a.kv_k, "K/V-cache output: operands not eligible for the tensor path");
  GN_REQUIRE(!a.red_add, "red_add: operands not eligible for the tensor path");
  GN_REQUIRE(!a.qkn_gamma, "qk-LayerNorm epilogue: operands not eligible for the tensor path");
  if (!a.force_simt) ++g_fallback_launches;   // CUDA-core GEMM on a handle that did not ask for the fp32 mode: shape / alignment cliff
  if (a.in_bf16) {
    if (a.fp16) return a.out_bf16 ? launch_simt<f16, f16>(a, stream) : launch_simt<f16, float>(a, stream);
    return a.out_bf16 ? launch_simt<bf16, bf16>(a, stream) : launch_simt<bf16, float>(a, stream);
  }
  return a.out_bf16 ? launch_simt<float, bf16>(a, stream) : launch_simt<float, float>(a, stream);
}

}  // namespace gn
