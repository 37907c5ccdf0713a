// Host-side orchestration of the GENIE forward / MaskGIT decode on one GPU + the C ABI
// (include/genie_b200.h).  All device work is enqueued on the caller's stream; the only host
// synchronisation points are the ones the header documents.
#include "kernels.cuh"
#include "../../include/genie_b200.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

using namespace gn;

namespace {

struct AttnW {
  void* qkv_w = nullptr;
  float* qkv_b = nullptr;
  void* proj_w = nullptr;
  float* proj_b = nullptr;
  float* norm_g = nullptr;
  float* norm_b = nullptr;
};
struct LayerW {
  AttnW attn[2];  // 0 spatial, 1 temporal
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  void* fc1_w = nullptr;
  float* fc1_b = nullptr;
  void* fc2_w = nullptr;
  float* fc2_b = nullptr;
  // folded LayerNorm (bf16, pre-LN): fp32 masters + W*diag(gamma) / column sums / folded bias
  float *qkv_s_raw = nullptr, *fc1_raw = nullptr;
  bf16 *qkv_s_f = nullptr, *fc1_f = nullptr;
  float *qkv_s_cs = nullptr, *qkv_s_bf = nullptr, *fc1_cs = nullptr, *fc1_bf = nullptr;
};

}  // namespace

static int g_live_handles[gn::kMaxDevices] = {};   // gn_model handles alive per device (persisting-L2 reset at the last one)

struct gn_model {
  gn_config cfg;
  int device = 0;
  int act_bf16 = 1;     // activations / weight matrices between kernels are 16-bit (bf16 or fp16)
  int fp16 = 0;         // ... and that 16-bit format is IEEE fp16 (GN_PREC_FP16) instead of bf16
  int o16() const { return act_bf16 ? (fp16 ? 2 : 1) : 0; }   // launch_prep output code
  int force_simt = 0;
  int tf32 = 0;         // tcgen05 kind::tf32 parity mode: every GEMM operand is pre-rounded to tf32 (RN)
  int hid = 0, C = 0;   // mlp hidden, readout width NV*V
  std::vector<LayerW> layers;
  float* pos = nullptr;         // [T,S,d]
  float* mask_embed = nullptr;  // [d]
  float* E = nullptr;           // [NV, V, d]
  void* out_w = nullptr;        // [C, d]
  float* out_b = nullptr;       // [C]
  std::set<std::string> have;
  std::vector<void*> owned;

  // chunk workspace of the ACTIVE lane (select_lane swaps these with lane_ws[])
  int64_t ws_tokens = 0;
  float* x = nullptr;     // [n, d] fp32 residual stream
  void* a = nullptr;      // [n, d] act
  void* big = nullptr;    // [n, max(3d, hid)] act (qkv / mlp hidden)
  void* o = nullptr;      // [n, d] act
  float* stats = nullptr; // [n, d/64, 2] row-statistics partials for the folded LayerNorm
  int qkn_epi = 0;        // qk-LayerNorm applied by the QKV GEMM epilogue (bf16, head_dim 64): the attention kernels then
                          // see already-normalised q / k and the tcgen05 spatial + temporal v2 kernels apply
  int fuse_readout = 1;   // temperature-0 decode: readout GEMM fused with softmax / argmax / confidence (readout_sample.cu)
  int red_epi = 0;        // residual GEMMs that need no 16-bit copy of the new stream update x through TMA reduce-add
  int tv2 = 0;            // temporal attention v2: head-major K/V caches written by the temporal QKV GEMM epilogue
  void* scr_k = nullptr;  // [chunk clips * S][H][T][hd] scratch K/V for tv2 when the persistent cache is off
  void* scr_v = nullptr;
  // Lanes: clips are independent, so the chunks of one MaskGIT step are dealt round-robin to `lanes` streams
  // (lane 0 = the caller's stream, the others library-owned), each with its own workspace.  The GPU then has two
  // independent kernel chains to schedule: the tail wave of one lane's persistent GEMM is filled by the other lane's
  // next kernel and HBM-bound kernels (temporal attention, LayerNorm, proj+residual) overlap tensor-bound ones.
  struct LaneWs {
    int64_t ws_tokens = 0;
    float* x = nullptr; void* a = nullptr; void* big = nullptr; void* o = nullptr; float* stats = nullptr;
    void* scr_k = nullptr; void* scr_v = nullptr;
  };
  static constexpr int kMaxLanes = 4;
  LaneWs lane_ws[kMaxLanes];
  int lane = 0;                               // active lane
  int lanes = 1;
  cudaStream_t lane_stream[kMaxLanes] = {};   // [0] unused (caller's stream)
  cudaEvent_t lane_fork = nullptr;
  cudaEvent_t lane_join[kMaxLanes] = {};
  int fold = 0;           // folded-LayerNorm path active (bf16, qk_norm = 0, cfg.fold_ln)
  bool fold_dirty = true;
  int64_t rows_cap = 0;
  float* rows = nullptr;  // [rows_cap, C] fp32 logits rows

  // temporal K/V cache [L][cache_B][S][T][d] act (all frames of one spatial position contiguous);
  // with tv2 the per-position block is head-major instead: [L][cache_B][S][H][T][hd]
  int cache_B = 0;
  void* kcache = nullptr;
  void* vcache = nullptr;

  // decode buffers for B clips
  int dec_B = 0;
  float* logits_frame = nullptr;   // [B*S, C]
  int32_t* samples_tmp = nullptr;  // [B*S]
  float* conf = nullptr;           // [B*S]
  uint8_t* unmasked = nullptr;     // [B*S]
  int32_t* prompt_scratch = nullptr;  // [B,T,S]
  int32_t* samples_scratch = nullptr; // [B*S]
  uint8_t* weight = nullptr;       // [B,T,S]
  int* flag = nullptr;             // device int
  float* noise_dev = nullptr;      // host-API staging
  int64_t noise_cap = 0;
  int32_t* tokens_dev = nullptr;
  int64_t tokens_cap = 0;

  double flops_executed = 0.0;
  double bytes_executed = 0.0;   // algorithmic HBM bytes of the launches (operands + outputs once per launch)

  // CUDA graphs of run_layers, keyed by (b0, nb, t0, Tact, use_cache); invalidated when a buffer is reallocated
  struct GraphEntry { cudaGraphExec_t exec; double flops; double bytes; unsigned long long launches, fallbacks; };
  std::map<std::vector<int>, GraphEntry> graphs;

  size_t esz() const { return act_bf16 ? 2 : 4; }
};

namespace {

int dev_alloc(gn_model* m, void** p, size_t bytes) {
  GN_CUDA_CHECK(cudaMalloc(p, bytes ? bytes : 16));
  m->owned.push_back(*p);
  return GN_OK;
}
void dev_free(gn_model* m, void* p) {
  if (!p) return;
  for (auto& q : m->owned)
    if (q == p) { q = nullptr; break; }
  cudaFree(p);
}

void drop_graphs(gn_model* m) {
  for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second.exec);
  m->graphs.clear();
}

// make lane `i`'s workspace the active one (host enqueue is sequential, so swapping the pointers is enough)
void select_lane(gn_model* m, int i) {
  if (i == m->lane) return;
  gn_model::LaneWs& cur = m->lane_ws[m->lane];
  cur.ws_tokens = m->ws_tokens; cur.x = m->x; cur.a = m->a; cur.big = m->big; cur.o = m->o; cur.stats = m->stats;
  cur.scr_k = m->scr_k; cur.scr_v = m->scr_v;
  const gn_model::LaneWs& nx = m->lane_ws[i];
  m->ws_tokens = nx.ws_tokens; m->x = nx.x; m->a = nx.a; m->big = nx.big; m->o = nx.o; m->stats = nx.stats;
  m->scr_k = nx.scr_k; m->scr_v = nx.scr_v;
  m->lane = i;
}

// lanes usable for a call on stream `st`: side streams are created on first use
int lanes_for(gn_model* m, cudaStream_t st) {
  if (m->lanes <= 1 || st == nullptr || st == cudaStreamLegacy || g_prof_on) return 1;
  if (!m->lane_fork) {
    GN_CUDA_CHECK(cudaEventCreateWithFlags(&m->lane_fork, cudaEventDisableTiming));
    for (int i = 1; i < m->lanes; ++i) {
      GN_CUDA_CHECK(cudaStreamCreateWithFlags(&m->lane_stream[i], cudaStreamNonBlocking));
      GN_CUDA_CHECK(cudaEventCreateWithFlags(&m->lane_join[i], cudaEventDisableTiming));
    }
  }
  return m->lanes;
}

int ensure_workspace(gn_model* m, int64_t n) {
  if (n <= m->ws_tokens) return GN_OK;
  drop_graphs(m);
  const int d = m->cfg.d_model;
  const int64_t wide = std::max<int64_t>(3 * d, m->hid);
  dev_free(m, m->x); dev_free(m, m->a); dev_free(m, m->big); dev_free(m, m->o); dev_free(m, m->stats);
  dev_free(m, m->scr_k); dev_free(m, m->scr_v);
  m->x = nullptr; m->a = m->big = m->o = m->scr_k = m->scr_v = nullptr; m->stats = nullptr;
  m->ws_tokens = 0;
  GN_PROPAGATE(dev_alloc(m, (void**)&m->x, (size_t)n * d * 4));
  GN_PROPAGATE(dev_alloc(m, &m->a, (size_t)n * d * m->esz()));
  GN_PROPAGATE(dev_alloc(m, &m->big, (size_t)n * wide * m->esz()));
  GN_PROPAGATE(dev_alloc(m, &m->o, (size_t)n * d * m->esz()));
  GN_PROPAGATE(dev_alloc(m, (void**)&m->stats, (size_t)n * (d / 64 + 1) * 2 * sizeof(float)));
  if (m->tv2) {   // dense (cache-less) calls are possible on any handle (compute_logits, forward)
    GN_PROPAGATE(dev_alloc(m, &m->scr_k, (size_t)n * d * m->esz()));
    GN_PROPAGATE(dev_alloc(m, &m->scr_v, (size_t)n * d * m->esz()));
  }
  m->ws_tokens = n;
  return GN_OK;
}
int ensure_rows(gn_model* m, int64_t r) {
  if (r <= m->rows_cap) return GN_OK;
  dev_free(m, m->rows);
  m->rows_cap = 0;
  GN_PROPAGATE(dev_alloc(m, (void**)&m->rows, (size_t)r * m->C * 4));
  m->rows_cap = r;
  return GN_OK;
}
int ensure_cache(gn_model* m, int B) {
  if (B <= m->cache_B) return GN_OK;
  drop_graphs(m);
  dev_free(m, m->kcache); dev_free(m, m->vcache);
  m->cache_B = 0;
  const size_t bytes = (size_t)m->cfg.num_layers * B * m->cfg.T * m->cfg.S * m->cfg.d_model * m->esz();
  GN_PROPAGATE(dev_alloc(m, &m->kcache, bytes));
  GN_PROPAGATE(dev_alloc(m, &m->vcache, bytes));
  m->cache_B = B;
  return GN_OK;
}
int ensure_decode(gn_model* m, int B) {
  if (B <= m->dec_B) return GN_OK;
  const int S = m->cfg.S, T = m->cfg.T;
  dev_free(m, m->logits_frame); dev_free(m, m->samples_tmp); dev_free(m, m->conf); dev_free(m, m->unmasked);
  dev_free(m, m->prompt_scratch); dev_free(m, m->samples_scratch); dev_free(m, m->weight);
  m->dec_B = 0;
  GN_PROPAGATE(dev_alloc(m, (void**)&m->logits_frame, (size_t)B * S * m->C * 4));
  GN_PROPAGATE(dev_alloc(m, (void**)&m->samples_tmp, (size_t)B * S * 4));
  GN_PROPAGATE(dev_alloc(m, (void**)&m->conf, (size_t)B * S * 4));
  GN_PROPAGATE(dev_alloc(m, (void**)&m->unmasked, (size_t)B * S));
  GN_PROPAGATE(dev_alloc(m, (void**)&m->prompt_scratch, (size_t)B * T * S * 4));
  GN_PROPAGATE(dev_alloc(m, (void**)&m->samples_scratch, (size_t)B * S * 4));
  GN_PROPAGATE(dev_alloc(m, (void**)&m->weight, (size_t)B * T * S));
  if (!m->flag) GN_PROPAGATE(dev_alloc(m, (void**)&m->flag, sizeof(int)));
  m->dec_B = B;
  return GN_OK;
}

int chunk_clips_for(const gn_model* m, int Tact) {
  const int ct = m->cfg.chunk_tokens > 0 ? m->cfg.chunk_tokens : 32768;
  int c = ct / (Tact * m->cfg.S);
  return c < 1 ? 1 : c;
}

struct KvOut {   // temporal QKV projection: K/V columns go straight to the (head-major) caches
  void* k = nullptr;
  void* v = nullptr;
  int t0 = 0, Tact = 0, clips = 0;
};

struct LnFold {
  const float* stats = nullptr;   // consumer: row-statistics partials
  int np = 0;
  const float* colsum = nullptr;
  float* stats_out = nullptr;     // producer (residual epilogue)
};

int linear(gn_model* m, const void* A, int64_t lda, const void* W, int K, const float* bias, const float* resid,
           void* out, int64_t ldo, void* out2, int M, int N, int epi, int out_bf16, cudaStream_t st,
           const LnFold* lf = nullptr, const KvOut* kv = nullptr, const AttnW* qkn = nullptr, bool red_add = false) {
  LinearArgs la{};
  la.red_add = red_add ? 1 : 0;
  if (qkn && m->qkn_epi) {   // q and k columns [0, 2d) get LayerNorm(head_dim) with the attention's shared affine
    la.qkn_gamma = qkn->norm_g; la.qkn_beta = qkn->norm_b; la.qkn_cols = 2 * m->cfg.d_model;
    la.qkn_hd = m->cfg.d_model / m->cfg.num_heads;
  }
  if (kv && kv->k) {
    la.kv_k = kv->k; la.kv_v = kv->v; la.kv_d = m->cfg.d_model; la.kv_hd = m->cfg.d_model / m->cfg.num_heads;
    la.kv_T = m->cfg.T; la.kv_S = m->cfg.S; la.kv_Tact = kv->Tact; la.kv_t0 = kv->t0; la.kv_clips = kv->clips;
  }
  if (lf) {
    la.ln_stats = lf->stats; la.ln_np = lf->np; la.ln_d = m->cfg.d_model; la.ln_colsum = lf->colsum;
    la.stats_out = lf->stats_out;
  }
  la.A = A; la.lda = lda; la.W = W; la.ldw = K; la.bias = bias; la.resid = resid; la.ldr = N;
  la.out = out; la.ldo = ldo; la.out2 = out2; la.ldo2 = N;
  la.M = M; la.N = N; la.K = K; la.epi = epi; la.in_bf16 = m->act_bf16; la.out_bf16 = out_bf16; la.fp16 = m->fp16;
  la.force_simt = m->force_simt;
  // The attention output and the MLP hidden are read by exactly one GEMM whose N = d spans only one or two column
  // tiles: load them evict_first so that the residual stream, the weights and the outputs being written stay in L2.
  // (Measured: -2 ms on the residual GEMMs.  Not for the LayerNorm output feeding QKV / fc1: their 6-8 column tiles
  // re-read every A tile and the hint makes those re-reads miss, +5 ms.)
  static const bool stream_hint = [] {
    const char* e = getenv("GENIE_B200_STREAM_HINT");
    return !(e && e[0] == '0');
  }();
  la.a_evict_first = (stream_hint && (A == m->o || A == m->big)) ? 1 : 0;
  la.round_out_tf32 = (m->tf32 && epi == EPI_GELU) ? 1 : 0;
  m->flops_executed += 2.0 * M * (double)N * K;
  {
    const double e = (double)m->esz();
    m->bytes_executed += (double)M * K * e + (double)N * K * e + (double)M * N * (out_bf16 ? 2.0 : 4.0) +
                         ((resid || red_add) ? (double)M * N * 4.0 : 0.0) + (out2 ? (double)M * N * 2.0 : 0.0);
  }
  return linear_forward(la, st);
}

// LayerNorm / cast pass (launch_prep) + its algorithmic bytes: fp32 row in, operand-format row out
int prep(gn_model* m, const float* gamma, const float* beta, int n, cudaStream_t st, int S, int Tact) {
  const int d = m->cfg.d_model;
  m->bytes_executed += (double)n * d * (4.0 + (double)m->esz());
  return launch_prep(m->x, m->a, m->o16(), gamma, beta, n, d, 1.f, S, Tact, -1, st, m->tf32);
}

// temporal QKV projection + causal attention over the frames of each spatial position (st_transformer.py:77-78,
// attention.py:36-61) for layer l: input rows `ain` [n, d], output m->o.
int temporal_block(gn_model* m, int l, const void* ain, AttnArgs aa, int b0, int nb, int t0, int Tact, bool use_cache,
                   cudaStream_t st) {
  const gn_config& c = m->cfg;
  const int d = c.d_model, S = c.S, T = c.T;
  const int n = nb * Tact * S;
  const int bf = m->act_bf16;
  const AttnW& w = m->layers[l].attn[1];
  void *kc = nullptr, *vc = nullptr;
  if (use_cache) {
    const size_t layer_stride = (size_t)m->cache_B * T * S * d * m->esz();
    const size_t clip_off = (size_t)b0 * T * S * d * m->esz();
    kc = (char*)m->kcache + l * layer_stride + clip_off;
    vc = (char*)m->vcache + l * layer_stride + clip_off;
  }
  if (m->tv2) {
    if (!use_cache) {
      GN_REQUIRE(t0 == 0 && Tact == T, "temporal attention without the K/V cache needs the full window");
      kc = m->scr_k;
      vc = m->scr_v;
    }
    KvOut kv;
    kv.k = kc; kv.v = vc; kv.t0 = t0; kv.Tact = Tact; kv.clips = nb;
    GN_PROPAGATE(linear(m, ain, d, w.qkv_w, d, w.qkv_b, nullptr, m->big, 3 * d, nullptr, n, 3 * d, EPI_STORE, 1, st,
                        nullptr, &kv, &w));
    GN_PROPAGATE(launch_temporal_attention_v2(aa, nb, S, T, t0, Tact, kc, vc, st));
  } else {
    GN_PROPAGATE(linear(m, ain, d, w.qkv_w, d, w.qkv_b, nullptr, m->big, 3 * d, nullptr, n, 3 * d, EPI_STORE, bf, st,
                        nullptr, nullptr, &w));
    GN_PROPAGATE(launch_temporal_attention(aa, nb, S, T, t0, Tact, kc, vc, c.generic_attention || !bf, st));
  }
  m->flops_executed += 4.0 * (t0 + Tact) * (double)d * n;
  // Q + output rows, and K/V of frames [0, t0 + Tact) of every (clip, position)
  m->bytes_executed += (2.0 * n * d + 2.0 * (double)nb * S * (t0 + Tact) * d) * (double)m->esz();
  return GN_OK;
}

// Runs the L ST blocks on `n = nb*Tact*S` compact rows already present in m->x.
// reference: genie/st_transformer.py:70-83 (STBlock.forward), :115-120.
int run_layers_body(gn_model* m, int b0, int nb, int t0, int Tact, bool use_cache, cudaStream_t st);

// L2 set-aside for the residual stream (see common.cuh): sized once per device, window = the chunk's rows of m->x.
struct L2WindowScope {
  explicit L2WindowScope(const gn_model* m, int64_t n_rows) {
    // per device: bytes available for persisting lines (0: unsupported / disabled, -1: not probed yet)
    static int64_t setaside_dev[kMaxDevices], max_window_dev[kMaxDevices];
    static bool probed[kMaxDevices] = {};
    const int di = current_device();
    int64_t& setaside = setaside_dev[di];
    int64_t& max_window_bytes = max_window_dev[di];
    if (!probed[di]) {
      probed[di] = true;
      setaside = 0;
      max_window_bytes = 0;
      // MB of L2 set aside for the residual stream (0 = off).  Measured on B200 (126 MB L2), round 2 (fp16 operands,
      // reduce-add residual epilogue; 3 runs each, same box): 0 / 8 / 16 / 24 / 32 / 40 / 48 MB = 1908 / 1934 / 1945 /
      // 1949-1958 / 1952 / 1943 / 1935 frames/s; 64 MB and more lose 8-15 % (the wide GEMMs then miss on their operands)
      const char* e = getenv("GENIE_B200_L2_PERSIST");
      const int64_t want = (int64_t)(e ? atoi(e) : 24) << 20;
      int dev = 0, max_persist = 0, max_window = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
      cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
      if (want > 0 && max_persist > 0 && max_window > 0 &&
          cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)std::min<int64_t>(want, max_persist)) ==
              cudaSuccess) {
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
        setaside = (int64_t)std::min<size_t>(got, (size_t)max_window);
        max_window_bytes = max_window;
      }
      cudaGetLastError();
    }
    if (setaside <= 0 || m->x == nullptr) return;
    // the window may not exceed cudaDevAttrMaxAccessPolicyWindowSize (128 MB on B200): larger chunks keep the policy
    // on their first rows only (a launch with a larger window fails with cudaErrorInvalidValue)
    // (a window over the first `setaside` bytes only, hitRatio 1, measured the same as a random fraction of all rows)
    const int64_t bytes = std::min<int64_t>(n_rows * m->cfg.d_model * 4, max_window_bytes);
    g_l2_window.base_ptr = m->x;
    g_l2_window.num_bytes = (size_t)bytes;
    g_l2_window.hitRatio = bytes <= setaside ? 1.0f : (float)setaside / (float)bytes;
    g_l2_window.hitProp = cudaAccessPropertyPersisting;
    g_l2_window.missProp = cudaAccessPropertyNormal;
  }
  ~L2WindowScope() { g_l2_window = cudaAccessPolicyWindow{}; }
};

int run_layers(gn_model* m, int b0, int nb, int t0, int Tact, bool use_cache, cudaStream_t st) {
  L2WindowScope l2(m, (int64_t)nb * Tact * m->cfg.S);
  return run_layers_body(m, b0, nb, t0, Tact, use_cache, st);
}

int run_layers_body(gn_model* m, int b0, int nb, int t0, int Tact, bool use_cache, cudaStream_t st) {
  const gn_config& c = m->cfg;
  const int d = c.d_model, S = c.S, T = c.T, H = c.num_heads, hd = d / H;
  const int n = nb * Tact * S;
  const int bf = m->act_bf16;
  const float scale = c.use_mup ? 8.0f / hd : 1.0f / sqrtf((float)hd);
  const int tf = m->tf32;
  if (m->fold) {
    // ---- bf16, pre-LN, LayerNorm folded into the QKV / fc1 epilogues: no stand-alone LayerNorm pass.
    // m->a always holds bf16(x) (emitted by the residual epilogues), m->stats the row statistics of x.
    if (m->fold_dirty) {
      for (int l = 0; l < c.num_layers; ++l) {
        LayerW& w = m->layers[l];
        const size_t nq = (size_t)3 * d * d, nf = (size_t)m->hid * d;
        if (!w.qkv_s_f) {
          GN_PROPAGATE(dev_alloc(m, (void**)&w.qkv_s_f, nq * 2));
          GN_PROPAGATE(dev_alloc(m, (void**)&w.qkv_s_cs, (size_t)3 * d * 4));
          GN_PROPAGATE(dev_alloc(m, (void**)&w.qkv_s_bf, (size_t)3 * d * 4));
          GN_PROPAGATE(dev_alloc(m, (void**)&w.fc1_f, nf * 2));
          GN_PROPAGATE(dev_alloc(m, (void**)&w.fc1_cs, (size_t)m->hid * 4));
          GN_PROPAGATE(dev_alloc(m, (void**)&w.fc1_bf, (size_t)m->hid * 4));
        }
        GN_PROPAGATE(launch_fold_ln(w.qkv_s_raw, w.ln1_g, w.ln1_b, w.attn[0].qkv_b, w.qkv_s_f, w.qkv_s_cs, w.qkv_s_bf,
                                    3 * d, d, st));
        GN_PROPAGATE(launch_fold_ln(w.fc1_raw, w.ln2_g, w.ln2_b, w.fc1_b, w.fc1_f, w.fc1_cs, w.fc1_bf, m->hid, d, st));
      }
      m->fold_dirty = false;
    }
    GN_PROPAGATE(launch_prep_stats(m->x, (bf16*)m->a, m->stats, n, d, st));
    int np = 1;
    const int np_d = d / resid_block_n(d, d, true);          // partials per row written by the proj epilogue
    const int np_h = d / resid_block_n(d, m->hid, true);     // ... and by the fc2 epilogue
    for (int l = 0; l < c.num_layers; ++l) {
      const LayerW& w = m->layers[l];
      LnFold lf{};
      lf.stats = m->stats; lf.np = np; lf.colsum = w.qkv_s_cs;
      GN_PROPAGATE(linear(m, m->a, d, w.qkv_s_f, d, w.qkv_s_bf, nullptr, m->big, 3 * d, nullptr, n, 3 * d, EPI_STORE, 1,
                          st, &lf));
      AttnArgs aa{};
      aa.qkv = m->big; aa.out = m->o; aa.act_bf16 = 1; aa.fp16 = m->fp16; aa.n_heads = H; aa.head_dim = hd; aa.scale = scale;
      GN_PROPAGATE(launch_spatial_attention(aa, nb * Tact, S, c.generic_attention, st));
      m->flops_executed += 4.0 * S * (double)d * n;
      m->bytes_executed += 4.0 * n * d * (double)m->esz();
      GN_PROPAGATE(linear(m, m->o, d, w.attn[0].proj_w, d, w.attn[0].proj_b, m->x, m->x, d, m->a, n, d, EPI_RESID, 0, st));
      GN_PROPAGATE(temporal_block(m, l, m->a, aa, b0, nb, t0, Tact, use_cache, st));
      LnFold ps{};
      ps.stats_out = m->stats;
      GN_PROPAGATE(linear(m, m->o, d, w.attn[1].proj_w, d, w.attn[1].proj_b, m->x, m->x, d, m->a, n, d, EPI_RESID, 0, st,
                          &ps));
      LnFold lf2{};
      lf2.stats = m->stats; lf2.np = np_d; lf2.colsum = w.fc1_cs;
      GN_PROPAGATE(linear(m, m->a, d, w.fc1_f, d, w.fc1_bf, nullptr, m->big, m->hid, nullptr, n, m->hid, EPI_GELU, 1, st,
                          &lf2));
      const bool last = l + 1 == c.num_layers;
      GN_PROPAGATE(linear(m, m->big, m->hid, w.fc2_w, m->hid, w.fc2_b, m->x, m->x, d, last ? nullptr : m->a, n, d,
                          EPI_RESID, 0, st, last ? nullptr : &ps));
      np = np_h;
    }
    return GN_OK;
  }
  const bool cp = bf || tf;  // GEMM A operands need a converted copy of the fp32 stream (bf16 cast / tf32 rounding)
  bool a_is_x = false;       // does m->a currently hold convert(x)?
  for (int l = 0; l < c.num_layers; ++l) {
    const LayerW& w = m->layers[l];
    // ---------------- spatial attention: x += proj(attn(qkv(norm1(x))))
    const void* ain;
    if (!c.qk_norm) {
      GN_PROPAGATE(prep(m, w.ln1_g, w.ln1_b, n, st, S, Tact));
      ain = m->a;
    } else if (cp) {
      if (!a_is_x) GN_PROPAGATE(prep(m, nullptr, nullptr, n, st, S, Tact));
      ain = m->a;
    } else {
      ain = m->x;
    }
    GN_PROPAGATE(linear(m, ain, d, w.attn[0].qkv_w, d, w.attn[0].qkv_b, nullptr, m->big, 3 * d, nullptr, n, 3 * d,
                        EPI_STORE, bf, st, nullptr, nullptr, &w.attn[0]));
    AttnArgs aa{};
    aa.qkv = m->big; aa.out = m->o; aa.act_bf16 = bf; aa.fp16 = m->fp16; aa.n_heads = H; aa.head_dim = hd; aa.scale = scale;
    aa.round_tf32 = tf;
    if (!m->qkn_epi) { aa.qk_gamma = w.attn[0].norm_g; aa.qk_beta = w.attn[0].norm_b; }
    GN_PROPAGATE(launch_spatial_attention(aa, nb * Tact, S, c.generic_attention || !bf, st));
    m->flops_executed += 4.0 * S * (double)d * n;
    m->bytes_executed += 4.0 * n * d * (double)m->esz();   // q, k, v in, o out
    // the temporal QKV GEMM reads the un-normalised stream: emit its bf16 copy from this epilogue
    GN_PROPAGATE(linear(m, m->o, d, w.attn[0].proj_w, d, w.attn[0].proj_b, m->x, m->x, d, bf ? m->a : nullptr, n, d,
                        EPI_RESID, 0, st));
    if (tf) GN_PROPAGATE(prep(m, nullptr, nullptr, n, st, S, Tact));
    // ---------------- temporal attention (no LayerNorm in front: st_transformer.py:78)
    if (!m->qkn_epi) { aa.qk_gamma = w.attn[1].norm_g; aa.qk_beta = w.attn[1].norm_b; }
    GN_PROPAGATE(temporal_block(m, l, cp ? m->a : (const void*)m->x, aa, b0, nb, t0, Tact, use_cache, st));
    const bool copy_t = bf && c.qk_norm;
    // x += proj(o).  Without a 16-bit copy to emit, the update is a TMA reduce-add of (acc + bias) into x: the residual
    // rows never travel into the SM (same fp32 sum, bit-identical), and the epilogue has no load -> modify -> store chain
    const bool red = m->red_epi && (bf || tf) && d % 64 == 0;
    if (red && !copy_t)
      GN_PROPAGATE(linear(m, m->o, d, w.attn[1].proj_w, d, w.attn[1].proj_b, nullptr, m->x, d, nullptr, n, d, EPI_STORE,
                          0, st, nullptr, nullptr, nullptr, true));
    else
    GN_PROPAGATE(linear(m, m->o, d, w.attn[1].proj_w, d, w.attn[1].proj_b, m->x, m->x, d, copy_t ? m->a : nullptr, n,
                        d, EPI_RESID, 0, st));
    // ---------------- MLP: x += fc2(gelu(fc1(norm2(x))))
    if (!c.qk_norm) {
      GN_PROPAGATE(prep(m, w.ln2_g, w.ln2_b, n, st, S, Tact));
      ain = m->a;
    } else if (cp) {
      if (tf) GN_PROPAGATE(prep(m, nullptr, nullptr, n, st, S, Tact));
      ain = m->a;
    } else {
      ain = m->x;
    }
    GN_PROPAGATE(linear(m, ain, d, w.fc1_w, d, w.fc1_b, nullptr, m->big, m->hid, nullptr, n, m->hid, EPI_GELU, bf, st));
    const bool copy_m = bf && c.qk_norm && (l + 1 < c.num_layers);
    if (red && !copy_m)
      GN_PROPAGATE(linear(m, m->big, m->hid, w.fc2_w, m->hid, w.fc2_b, nullptr, m->x, d, nullptr, n, d, EPI_STORE, 0, st,
                          nullptr, nullptr, nullptr, true));
    else
    GN_PROPAGATE(linear(m, m->big, m->hid, w.fc2_w, m->hid, w.fc2_b, m->x, m->x, d, copy_m ? m->a : nullptr, n, d,
                        EPI_RESID, 0, st));
    a_is_x = copy_m;
  }
  return GN_OK;
}

// run_layers through a CUDA graph: the ~10 launches x L layers of one chunk are captured the first time a
// (b0, nb, t0, Tact, cache) shape is seen on a capturable stream and replayed afterwards (all pointers they use are
// model-owned and stable; reallocation drops the graphs).
int run_layers_graphed(gn_model* m, int b0, int nb, int t0, int Tact, bool use_cache, cudaStream_t st) {
  if (!m->cfg.cuda_graphs || st == nullptr || st == cudaStreamLegacy || g_prof_on || (m->fold && m->fold_dirty))
    return run_layers(m, b0, nb, t0, Tact, use_cache, st);
  const std::vector<int> key = {b0, nb, t0, Tact, use_cache ? 1 : 0, m->lane};
  auto it = m->graphs.find(key);
  if (it == m->graphs.end()) {
    const double f0 = m->flops_executed, y0 = m->bytes_executed;
    const unsigned long long l0 = g_launch_count, fb0 = g_fallback_launches;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      cudaGetLastError();
      return run_layers(m, b0, nb, t0, Tact, use_cache, st);
    }
    const int rc = run_layers(m, b0, nb, t0, Tact, use_cache, st);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != GN_OK || ce != cudaSuccess || graph == nullptr) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      if (rc != GN_OK) return rc;
      m->flops_executed = f0;
      m->bytes_executed = y0;
      return run_layers(m, b0, nb, t0, Tact, use_cache, st);   // capture not possible here: run eagerly
    }
    gn_model::GraphEntry e{};
    e.flops = m->flops_executed - f0;
    e.bytes = m->bytes_executed - y0;
    e.launches = g_launch_count - l0;
    e.fallbacks = g_fallback_launches - fb0;
    const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
    cudaGraphDestroy(graph);
    m->flops_executed = f0;
    m->bytes_executed = y0;
    g_launch_count = l0;
    g_fallback_launches = fb0;
    if (ie != cudaSuccess) {
      cudaGetLastError();
      return run_layers(m, b0, nb, t0, Tact, use_cache, st);
    }
    it = m->graphs.emplace(key, e).first;
  }
  GN_CUDA_CHECK(cudaGraphLaunch(it->second.exec, st));
  m->flops_executed += it->second.flops;
  m->bytes_executed += it->second.bytes;
  g_launch_count += it->second.launches;
  g_fallback_launches += it->second.fallbacks;
  return GN_OK;
}

// embed (+pos) clips [b0, b0+nb), frames [t0, t0+Tact) of ids [B,T,S] into m->x, then the L blocks.
int forward_chunk(gn_model* m, const int32_t* ids, int b0, int nb, int t0, int Tact, bool use_cache, cudaStream_t st) {
  const gn_config& c = m->cfg;
  GN_PROPAGATE(ensure_workspace(m, (int64_t)nb * Tact * c.S));
  GN_PROPAGATE(launch_embed(ids + (int64_t)b0 * c.T * c.S, m->E, m->mask_embed, m->pos, m->x, nb, c.T, c.S, t0, Tact,
                            c.d_model, c.factored_vocab_size, c.num_factored_vocabs, c.image_vocab_size, st));
  return run_layers_graphed(m, b0, nb, t0, Tact, use_cache, st);
}

// readout of rows (optionally only local frame `tsel`) of m->x into out_rows [R, C] fp32
// reference: st_mask_git.py:262 (+ FixedMuReadout :316-323: input scaled by 256/d under muP)
int readout(gn_model* m, int nb, int Tact, int tsel, float* out_rows, cudaStream_t st) {
  const gn_config& c = m->cfg;
  const int R = tsel >= 0 ? nb * c.S : nb * Tact * c.S;
  const float mult = c.use_mup ? 256.0f / c.d_model : 1.0f;
  const void* ain;
  if (!m->act_bf16 && !m->tf32 && tsel < 0 && mult == 1.0f) {
    ain = m->x;
  } else {
    GN_PROPAGATE(launch_prep(m->x, m->a, m->o16(), nullptr, nullptr, R, c.d_model, mult, c.S, Tact, tsel, st,
                             m->tf32));
    ain = m->a;
  }
  return linear(m, ain, c.d_model, m->out_w, c.d_model, m->out_b, nullptr, out_rows, m->C, nullptr, R, m->C, EPI_STORE,
                0, st);
}

// readout of local frame `tsel` fused with the factored softmax / argmax / confidence (readout_sample.cu): ids and
// confidences of clips [b0, b0 + nb) go straight to samples / conf; the logits rows are written only when `logits_rows`
// is non-null (step-0 logits of maskgit_generate, CE of evaluate).  reference: st_mask_git.py:262 + :171-190.
int readout_and_sample(gn_model* m, int nb, int Tact, int tsel, float* logits_rows, int32_t* samples, float* conf,
                       cudaStream_t st) {
  const gn_config& c = m->cfg;
  const int R = nb * c.S;
  const float mult = c.use_mup ? 256.0f / c.d_model : 1.0f;
  GN_PROPAGATE(launch_prep(m->x, m->a, m->o16(), nullptr, nullptr, R, c.d_model, mult, c.S, Tact, tsel, st, 0));
  m->flops_executed += 2.0 * R * (double)m->C * c.d_model;
  m->bytes_executed += (double)R * c.d_model * (4.0 + 2.0 * m->esz()) + (double)m->C * c.d_model * m->esz() +
                       (logits_rows ? (double)R * m->C * 4.0 : 0.0) + 8.0 * R;
  return launch_readout_sample(m->a, m->out_w, m->out_b, logits_rows, samples, conf, R, c.d_model, c.num_factored_vocabs,
                               m->fp16, st);
}

__global__ void fill_i32_kernel(int32_t* p, int64_t clip_stride, int64_t off, int64_t per_clip, int B, int32_t v) {
  const int64_t total = (int64_t)B * per_clip;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    p[(i / per_clip) * clip_stride + off + (i % per_clip)] = v;
  }
}
int fill_frames(int32_t* tokens, int B, int T, int S, int t_from, int32_t v, cudaStream_t st) {
  if (t_from >= T) return GN_OK;
  const int64_t per = (int64_t)(T - t_from) * S;
  const int grid = (int)std::min<int64_t>(ceil_div64((int64_t)B * per, 256), 1184);
  fill_i32_kernel<<<grid, 256, 0, st>>>(tokens, (int64_t)T * S, (int64_t)t_from * S, per, B, v);
  GN_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return GN_OK;
}

__global__ void relevant_weight_kernel(const int32_t* __restrict__ ids, uint8_t* __restrict__ w, int64_t total, int TS,
                                       int S, int mask_id) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)((i % TS) / S);
    w[i] = (t > 0 && ids[i] == mask_id) ? 1 : 0;
  }
}

// One maskgit_generate call (st_mask_git.py:123-229).  `recompute_from`: first frame whose hidden state /
// temporal K,V must be (re)computed at step 0 when the K/V cache is on (0 for a stand-alone call).
// CE hooks: when `ce_targets` != nullptr the step-0 logits are scored against it (evaluate.py:177).
int maskgit_impl(gn_model* m, int32_t* prompt, int B, int out_t, int steps, int unmask_mode, const float* noise,
                 const float* uniform, int32_t* samples, float* logits0, int logits0_Tout, int logits0_slot,
                 int recompute_from, const int32_t* ce_targets, double* acc, cudaStream_t st) {
  const gn_config& c = m->cfg;
  const int S = c.S, T = c.T;
  GN_PROPAGATE(ensure_decode(m, B));
  const bool cache = c.kv_cache != 0;
  if (cache) GN_PROPAGATE(ensure_cache(m, B));
  GN_CUDA_CHECK(cudaMemsetAsync(m->unmasked, 0, (size_t)B * S, st));
  for (int step = 0; step < steps; ++step) {
    int t0 = 0, Tact = T;
    if (cache) {
      t0 = step == 0 ? recompute_from : out_t;
      Tact = out_t + 1 - t0;
    }
    const int nl = lanes_for(m, st);
    if (nl < 0) return nl;
    const int cc = std::max(1, std::min(chunk_clips_for(m, Tact), (B + nl - 1) / nl));
    const int n_chunks = (B + cc - 1) / cc;
    const int used = std::min(nl, n_chunks);
    if (used > 1) {   // fork: the side lanes start after everything already enqueued on the caller's stream
      GN_CUDA_CHECK(cudaEventRecord(m->lane_fork, st));
      for (int i = 1; i < used; ++i) GN_CUDA_CHECK(cudaStreamWaitEvent(m->lane_stream[i], m->lane_fork, 0));
    }
    // temperature 0 on the 16-bit tensor path: readout GEMM fused with the decode math; the logits rows are only
    // materialised when this step's logits are consumed (step-0 return value / CE)
    const bool need_logits = step == 0 && (logits0 != nullptr || ce_targets != nullptr);
    const bool fused = m->fuse_readout && uniform == nullptr && m->out_b != nullptr &&
                       readout_sample_supported(c.factored_vocab_size, c.d_model, m->act_bf16 && !m->force_simt);
    int rc = GN_OK;
    for (int b0 = 0, ci = 0; b0 < B && rc == GN_OK; b0 += cc, ++ci) {
      const int nb = std::min(cc, B - b0);
      const int li = ci % used;
      cudaStream_t ls = li == 0 ? st : m->lane_stream[li];
      select_lane(m, li);
      rc = forward_chunk(m, prompt, b0, nb, t0, Tact, cache, ls);
      if (rc != GN_OK) break;
      if (fused)
        rc = readout_and_sample(m, nb, Tact, out_t - t0, need_logits ? m->logits_frame + (int64_t)b0 * S * m->C : nullptr,
                                m->samples_tmp + (int64_t)b0 * S, m->conf + (int64_t)b0 * S, ls);
      else
        rc = readout(m, nb, Tact, out_t - t0, m->logits_frame + (int64_t)b0 * S * m->C, ls);
    }
    select_lane(m, 0);
    if (used > 1) {   // join (also on the error path: the caller's stream must not run ahead of the side lanes)
      for (int i = 1; i < used; ++i) {
        GN_CUDA_CHECK(cudaEventRecord(m->lane_join[i], m->lane_stream[i]));
        GN_CUDA_CHECK(cudaStreamWaitEvent(st, m->lane_join[i], 0));
      }
    }
    GN_PROPAGATE(rc);
    if (step == 0) {
      if (logits0) GN_PROPAGATE(launch_logits_transpose(m->logits_frame, logits0, B, 1, S, m->C, logits0_Tout,
                                                        logits0_slot, st));
      if (ce_targets)
        GN_PROPAGATE(launch_ce(m->logits_frame, ce_targets, (int64_t)T * S, S, B * S, c.factored_vocab_size,
                               c.num_factored_vocabs, nullptr, acc, st));
    }
    // uniform [steps, B, S, NV]: Categorical draw of this step (st_mask_git.py:182-187); nullptr = argmax
    const float* un = uniform ? uniform + (int64_t)step * B * S * c.num_factored_vocabs : nullptr;
    if (!fused)
      GN_PROPAGATE(launch_sample(m->logits_frame, B * S, c.factored_vocab_size, c.num_factored_vocabs, un,
                                 m->samples_tmp, m->conf, st));
    const bool last = step == steps - 1;
    const int n_mask = last ? 0 : (int)std::ceil(std::cos((step + 1.0) / steps * M_PI / 2.0) * S);
    const float* cf = nullptr;
    if (!last) cf = unmask_mode == GN_UNMASK_GREEDY ? m->conf : noise + (int64_t)step * B * S;
    GN_PROPAGATE(launch_remask(prompt + (int64_t)out_t * S, (int64_t)T * S, m->samples_tmp, cf, m->unmasked, samples, B,
                               S, n_mask, last ? 1 : 0, c.image_vocab_size, st));
  }
  return GN_OK;
}

int check_generate_args(gn_model* m, int B, int steps, float temperature, int unmask_mode, const float* noise,
                        const float* uniform) {
  GN_REQUIRE(m != nullptr, "null model handle");
  GN_REQUIRE(B > 0, "batch must be positive (got %d)", B);
  GN_REQUIRE(steps >= 1, "maskgit_steps must be >= 1 (got %d)", steps);
  GN_REQUIRE(temperature <= 1e-8f || uniform != nullptr,
             "temperature > 0 (Categorical sampling, st_mask_git.py:182-187) needs the caller's uniform tensor "
             "[.., steps, B, S, num_factored_vocabs]");
  GN_REQUIRE(unmask_mode == GN_UNMASK_RANDOM || unmask_mode == GN_UNMASK_GREEDY,
             "Expected `unmask_mode` to be one of ['greedy', 'random'], got %d", unmask_mode);
  GN_REQUIRE(steps == 1 || unmask_mode == GN_UNMASK_GREEDY || noise != nullptr,
             "unmask_mode='random' with maskgit_steps > 1 needs the caller's noise tensor [steps-1, B, S]");
  return gn_model_check_weights(m);
}

int sync_check_flag(gn_model* m, int out_t, cudaStream_t st) {
  int h = 0;
  GN_CUDA_CHECK(cudaMemcpyAsync(&h, m->flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  GN_CUDA_CHECK(cudaStreamSynchronize(st));
  GN_REQUIRE(h == 0, "when generating z%d, frames %d and later must be masked", out_t, out_t);
  return GN_OK;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int gn_version(void) { return GN_ABI_VERSION; }
const char* gn_last_error(void) { return gn::last_error(); }
uint64_t gn_kernel_launches(void) { return gn::g_launch_count; }
uint64_t gn_fallback_launches(void) { return gn::g_fallback_launches; }

int gn_model_create(gn_model** out, const gn_config* cfg, int device) {
  GN_REQUIRE(out && cfg, "gn_model_create: null argument");
  *out = nullptr;
  GN_REQUIRE(cfg->num_layers > 0 && cfg->num_heads > 0 && cfg->d_model > 0, "invalid model dims");
  GN_REQUIRE(cfg->d_model % cfg->num_heads == 0, "d_model %% num_heads != 0");
  GN_REQUIRE(cfg->d_model % 8 == 0, "d_model must be a multiple of 8");
  GN_REQUIRE(cfg->T > 0 && cfg->S > 0, "invalid T/S");
  {
    int h = (int)std::lround(std::sqrt((double)cfg->S));
    GN_REQUIRE(h * h == cfg->S, "Expected S to be square");
  }
  GN_REQUIRE(cfg->num_factored_vocabs >= 1 && cfg->factored_vocab_size >= 2, "invalid vocab factorisation");
  {
    int64_t p = 1;
    for (int i = 0; i < cfg->num_factored_vocabs; ++i) p *= cfg->factored_vocab_size;
    GN_REQUIRE(p == cfg->image_vocab_size, "factored_vocab_size ** num_factored_vocabs != image_vocab_size");
  }
  GN_REQUIRE(cfg->precision >= GN_PREC_BF16 && cfg->precision <= GN_PREC_FP16, "unknown precision %d", cfg->precision);
  int ndev = 0;
  GN_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  GN_REQUIRE(device >= 0 && device < ndev, "device %d not available (%d visible)", device, ndev);
  cudaDeviceProp prop;
  GN_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  GN_REQUIRE(prop.major == 10, "libgenie_b200 is built for sm_100a only (device %d is sm_%d%d)", device, prop.major,
             prop.minor);
  gn_model* m = new (std::nothrow) gn_model();
  GN_REQUIRE(m, "out of host memory");
  m->cfg = *cfg;
  m->device = device;
  m->act_bf16 = cfg->precision == GN_PREC_BF16 || cfg->precision == GN_PREC_FP16;
  m->fp16 = cfg->precision == GN_PREC_FP16;
  m->force_simt = cfg->precision == GN_PREC_FP32;
  m->tf32 = cfg->precision == GN_PREC_TF32;
  m->fold = (m->act_bf16 && !m->fp16 && !cfg->qk_norm && cfg->fold_ln && cfg->d_model % 64 == 0) ? 1 : 0;
  {
    const char* e = getenv("GENIE_B200_TEMPORAL_V2");
    const bool on = !(e && (e[0] == '0' || e[0] == 'n' || e[0] == 'N'));
    AttnArgs probe{};
    probe.act_bf16 = m->act_bf16; probe.fp16 = m->fp16; probe.n_heads = cfg->num_heads; probe.head_dim = cfg->d_model / cfg->num_heads;
    const char* q = getenv("GENIE_B200_QKN_EPI");   // 0: keep qk-LayerNorm inside the (mma.sync) attention kernels
    const bool qon = !(q && (q[0] == '0' || q[0] == 'n' || q[0] == 'N'));
    m->qkn_epi = (qon && cfg->qk_norm && m->act_bf16 && !cfg->generic_attention &&
                  ((probe.head_dim == 64 && (3 * cfg->d_model) % 128 == 0) ||
                   (probe.head_dim == 32 && (3 * cfg->d_model) % 64 == 0))) ? 1 : 0;
    m->tv2 = (on && (!cfg->qk_norm || m->qkn_epi) && !cfg->generic_attention &&
              temporal_v2_supported(probe, cfg->S, cfg->T)) ? 1 : 0;
  }
  {
    // A/B switch, default on: measured on B200 (GENIE_138M, 64 clips): residual GEMMs 82.2 -> 78.7 ms per step,
    // 1893 -> 1925 frames/s, results bit-identical (tests/test_gpu_model.py)
    const char* e = getenv("GENIE_B200_RED_EPI");
    m->red_epi = e ? (e[0] != '0') : 1;
    const char* f = getenv("GENIE_B200_FUSED_READOUT");   // 0: readout GEMM + sample_kernel (two launches, logits via HBM)
    m->fuse_readout = f ? (f[0] != '0') : 1;
  }
  m->lanes = cfg->lanes <= 0 ? 1 : std::min<int>(cfg->lanes, gn_model::kMaxLanes);   // measured neutral on B200: off by default
  m->hid = (int)(cfg->d_model * cfg->mlp_ratio);
  m->C = cfg->num_factored_vocabs * cfg->factored_vocab_size;
  m->layers.resize(cfg->num_layers);
  ++g_live_handles[device < kMaxDevices ? device : 0];
  *out = m;
  return GN_OK;
}

void gn_model_destroy(gn_model* m) {
  if (!m) return;
  DeviceGuard g(m->device);
  drop_graphs(m);
  for (int i = 1; i < gn_model::kMaxLanes; ++i) {
    if (m->lane_stream[i]) { cudaStreamSynchronize(m->lane_stream[i]); cudaStreamDestroy(m->lane_stream[i]); }
    if (m->lane_join[i]) cudaEventDestroy(m->lane_join[i]);
  }
  if (m->lane_fork) cudaEventDestroy(m->lane_fork);
  for (void* p : m->owned)
    if (p) cudaFree(p);
  // lines of the (now freed) residual stream leave the L2 set-aside - only when no other handle on this device is
  // alive: the reset evicts EVERY persisting line of the context (other handles' windows, PyTorch kernels' policies)
  if (--g_live_handles[m->device < kMaxDevices ? m->device : 0] <= 0) cudaCtxResetPersistingL2Cache();
  cudaGetLastError();
  delete m;
}

int gn_model_set_weight(gn_model* m, const char* key, const float* src, const int64_t* shape, int ndim, void* stream) {
  GN_REQUIRE(m && key && src && shape, "gn_model_set_weight: null argument");
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  const gn_config& c = m->cfg;
  const int d = c.d_model, hd = d / c.num_heads;
  m->fold_dirty = true;
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= shape[i];
  std::string k(key);

  auto want = [&](int64_t n) -> int {
    GN_REQUIRE(numel == n, "weight %s: expected %lld elements, got %lld", key, (long long)n, (long long)numel);
    return GN_OK;
  };
  auto put_f32 = [&](float** dst, int64_t n) -> int {
    GN_PROPAGATE(want(n));
    if (!*dst) GN_PROPAGATE(dev_alloc(m, (void**)dst, (size_t)n * 4));
    GN_CUDA_CHECK(cudaMemcpyAsync(*dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    m->have.insert(k);
    return GN_OK;
  };
  auto put_mat = [&](void** dst, int64_t n) -> int {
    GN_PROPAGATE(want(n));
    if (!*dst) GN_PROPAGATE(dev_alloc(m, dst, (size_t)n * m->esz()));
    if (m->act_bf16) GN_PROPAGATE(launch_cast_h16(src, *dst, m->fp16, n, st));
    else if (m->tf32) GN_PROPAGATE(launch_round_tf32(src, (float*)*dst, n, st));
    else GN_CUDA_CHECK(cudaMemcpyAsync(*dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    m->have.insert(k);
    return GN_OK;
  };

  if (k == "pos_embed_TSC") return put_f32(&m->pos, (int64_t)c.T * c.S * d);
  if (k == "token_embed.mask_token_embed") return put_f32(&m->mask_embed, d);
  if (k.rfind("token_embed.factored_embeds.", 0) == 0) {
    int i = -1;
    if (sscanf(key, "token_embed.factored_embeds.%d.weight", &i) != 1 || i < 0 || i >= c.num_factored_vocabs) {
      set_error("unknown weight key %s", key);
      return GN_ERR_INVALID;
    }
    const int64_t n = (int64_t)c.factored_vocab_size * d;
    GN_PROPAGATE(want(n));
    if (!m->E) GN_PROPAGATE(dev_alloc(m, (void**)&m->E, (size_t)c.num_factored_vocabs * n * 4));
    GN_CUDA_CHECK(cudaMemcpyAsync(m->E + (int64_t)i * n, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    m->have.insert(k);
    return GN_OK;
  }
  if (k == "out_x_proj.weight") return put_mat(&m->out_w, (int64_t)m->C * d);
  if (k == "out_x_proj.bias") return put_f32(&m->out_b, m->C);
  int l = -1;
  char rest[128] = "";
  if (sscanf(key, "decoder.layers.%d.%127s", &l, rest) == 2 && l >= 0 && l < c.num_layers) {
    LayerW& w = m->layers[l];
    std::string r(rest);
    for (int which = 0; which < 2; ++which) {
      const std::string p = which == 0 ? "spatial_attn." : "temporal_attn.";
      if (r.rfind(p, 0) != 0) continue;
      const std::string t = r.substr(p.size());
      AttnW& a = w.attn[which];
      if (t == "qkv.weight") {
        if (which == 0 && m->fold) { GN_PROPAGATE(put_f32(&w.qkv_s_raw, (int64_t)3 * d * d)); m->have.erase(k); }
        return put_mat(&a.qkv_w, (int64_t)3 * d * d);
      }
      if (t == "qkv.bias") return put_f32(&a.qkv_b, 3 * d);
      if (t == "proj.weight") return put_mat(&a.proj_w, (int64_t)d * d);
      if (t == "proj.bias") return put_f32(&a.proj_b, d);
      if (t == "norm.weight") return put_f32(&a.norm_g, hd);
      if (t == "norm.bias") return put_f32(&a.norm_b, hd);
    }
    if (r == "norm1.weight") return put_f32(&w.ln1_g, d);
    if (r == "norm1.bias") return put_f32(&w.ln1_b, d);
    if (r == "norm2.weight") return put_f32(&w.ln2_g, d);
    if (r == "norm2.bias") return put_f32(&w.ln2_b, d);
    if (r == "mlp.fc1.weight") {
      if (m->fold) { GN_PROPAGATE(put_f32(&w.fc1_raw, (int64_t)m->hid * d)); m->have.erase(k); }
      return put_mat(&w.fc1_w, (int64_t)m->hid * d);
    }
    if (r == "mlp.fc1.bias") return put_f32(&w.fc1_b, m->hid);
    if (r == "mlp.fc2.weight") return put_mat(&w.fc2_w, (int64_t)d * m->hid);
    if (r == "mlp.fc2.bias") return put_f32(&w.fc2_b, d);
  }
  set_error("unknown weight key %s", key);
  return GN_ERR_INVALID;
}

int gn_model_check_weights(gn_model* m) {
  GN_REQUIRE(m, "null model handle");
  const gn_config& c = m->cfg;
  std::vector<std::string> need = {"pos_embed_TSC", "token_embed.mask_token_embed", "out_x_proj.weight",
                                   "out_x_proj.bias"};
  for (int i = 0; i < c.num_factored_vocabs; ++i)
    need.push_back("token_embed.factored_embeds." + std::to_string(i) + ".weight");
  for (int l = 0; l < c.num_layers; ++l) {
    const std::string p = "decoder.layers." + std::to_string(l) + ".";
    for (const char* a : {"spatial_attn.", "temporal_attn."}) {
      need.push_back(p + a + "qkv.weight");
      need.push_back(p + a + "proj.weight");
      if (c.qkv_bias) need.push_back(p + a + "qkv.bias");
      if (c.proj_bias) need.push_back(p + a + "proj.bias");
      if (c.qk_norm) { need.push_back(p + a + "norm.weight"); need.push_back(p + a + "norm.bias"); }
    }
    if (!c.qk_norm)
      for (const char* nm : {"norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias"}) need.push_back(p + nm);
    need.push_back(p + "mlp.fc1.weight");
    need.push_back(p + "mlp.fc2.weight");
    if (c.mlp_bias) { need.push_back(p + "mlp.fc1.bias"); need.push_back(p + "mlp.fc2.bias"); }
  }
  for (const auto& k : need) {
    if (!m->have.count(k)) {
      set_error("missing weight %s", k.c_str());
      return GN_ERR_STATE;
    }
  }
  return GN_OK;
}

int gn_decoder_forward(gn_model* m, const float* x, float* y, int B, void* stream) {
  GN_REQUIRE(m && x && y && B > 0, "gn_decoder_forward: invalid argument");
  GN_PROPAGATE(gn_model_check_weights(m));
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  const gn_config& c = m->cfg;
  const int cc = chunk_clips_for(m, c.T);
  const int64_t per_clip = (int64_t)c.T * c.S * c.d_model;
  for (int b0 = 0; b0 < B; b0 += cc) {
    const int nb = std::min(cc, B - b0);
    GN_PROPAGATE(ensure_workspace(m, (int64_t)nb * c.T * c.S));
    GN_CUDA_CHECK(cudaMemcpyAsync(m->x, x + b0 * per_clip, (size_t)nb * per_clip * 4, cudaMemcpyDeviceToDevice, st));
    GN_PROPAGATE(run_layers(m, b0, nb, 0, c.T, false, st));
    GN_CUDA_CHECK(cudaMemcpyAsync(y + b0 * per_clip, m->x, (size_t)nb * per_clip * 4, cudaMemcpyDeviceToDevice, st));
  }
  return GN_OK;
}

int gn_attention_forward(gn_model* m, int layer, int which, const float* x, float* y, int n_seq, int n_tok, int causal,
                         void* stream) {
  GN_REQUIRE(m && x && y && n_seq > 0 && n_tok > 0, "gn_attention_forward: invalid argument");
  GN_REQUIRE(layer >= 0 && layer < m->cfg.num_layers && (which == 0 || which == 1), "bad layer/which");
  GN_PROPAGATE(gn_model_check_weights(m));
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  const gn_config& c = m->cfg;
  const int d = c.d_model, hd = d / c.num_heads, n = n_seq * n_tok;
  const int bf = m->act_bf16;
  const AttnW& w = m->layers[layer].attn[which];
  GN_PROPAGATE(ensure_workspace(m, n));
  const void* ain = x;
  if (bf || m->tf32) {
    GN_PROPAGATE(launch_prep(x, m->a, m->o16(), nullptr, nullptr, n, d, 1.f, 1, 1, -1, st, m->tf32));
    ain = m->a;
  }
  GN_PROPAGATE(linear(m, ain, d, w.qkv_w, d, w.qkv_b, nullptr, m->big, 3 * d, nullptr, n, 3 * d, EPI_STORE, bf, st));
  AttnArgs aa{};
  aa.qkv = m->big; aa.out = m->o; aa.act_bf16 = bf; aa.fp16 = m->fp16; aa.n_heads = c.num_heads; aa.head_dim = hd;
  aa.scale = c.use_mup ? 8.0f / hd : 1.0f / sqrtf((float)hd);
  aa.qk_gamma = w.norm_g; aa.qk_beta = w.norm_b; aa.round_tf32 = m->tf32;
  GN_PROPAGATE(launch_generic_attention(aa, n_seq, n_tok, causal, st));
  return linear(m, m->o, d, w.proj_w, d, w.proj_b, nullptr, y, d, nullptr, n, d, EPI_STORE, 0, st);
}

int gn_compute_logits(gn_model* m, const int32_t* ids, int B, float* logits, void* stream) {
  GN_REQUIRE(m && ids && logits && B > 0, "gn_compute_logits: invalid argument");
  GN_PROPAGATE(gn_model_check_weights(m));
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  const gn_config& c = m->cfg;
  const int cc = chunk_clips_for(m, c.T);
  for (int b0 = 0; b0 < B; b0 += cc) {
    const int nb = std::min(cc, B - b0);
    GN_PROPAGATE(forward_chunk(m, ids, b0, nb, 0, c.T, false, st));
    GN_PROPAGATE(ensure_rows(m, (int64_t)nb * c.T * c.S));
    GN_PROPAGATE(readout(m, nb, c.T, -1, m->rows, st));
    GN_PROPAGATE(launch_logits_transpose(m->rows, logits + (int64_t)b0 * m->C * c.T * c.S, nb, c.T, c.S, m->C, c.T, 0,
                                         st));
  }
  return GN_OK;
}

int gn_maskgit_generate(gn_model* m, int32_t* prompt, int B, int out_t, int steps, float temperature, int unmask_mode,
                        const float* noise, const float* uniform, int32_t* samples, float* logits0, void* stream) {
  GN_PROPAGATE(check_generate_args(m, B, steps, temperature, unmask_mode, noise, uniform));
  if (temperature <= 1e-8f) uniform = nullptr;
  GN_REQUIRE(prompt && samples, "gn_maskgit_generate: null buffer");
  GN_REQUIRE(out_t > 0, "maskgit_generate requires out_t > 0");
  GN_REQUIRE(out_t < m->cfg.T, "out_t %d out of range (T=%d)", out_t, m->cfg.T);
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  GN_PROPAGATE(ensure_decode(m, B));
  GN_CUDA_CHECK(cudaMemsetAsync(m->flag, 0, sizeof(int), st));
  GN_PROPAGATE(launch_check_masked(prompt, B, m->cfg.T, m->cfg.S, out_t, m->cfg.image_vocab_size, m->flag, st));
  // The reference asserts before doing any work (st_mask_git.py:155); we must not mutate the prompt if
  // the precondition fails, so this one check is synchronous, like the reference's.
  GN_PROPAGATE(sync_check_flag(m, out_t, st));
  return maskgit_impl(m, prompt, B, out_t, steps, unmask_mode, noise, uniform, samples, logits0, 1, 0, 0, nullptr,
                      nullptr, st);
}

int gn_generate(gn_model* m, int32_t* tokens, int B, int t_prompt, int steps, float temperature, int unmask_mode,
                const float* noise, const float* uniform, float* logits0, void* stream) {
  GN_PROPAGATE(check_generate_args(m, B, steps, temperature, unmask_mode, noise, uniform));
  if (temperature <= 1e-8f) uniform = nullptr;
  GN_REQUIRE(tokens, "gn_generate: null buffer");
  const gn_config& c = m->cfg;
  GN_REQUIRE(t_prompt >= 1 && t_prompt <= c.T, "num_prompt_frames %d out of range [1, %d]", t_prompt, c.T);
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  GN_PROPAGATE(ensure_decode(m, B));
  GN_PROPAGATE(fill_frames(tokens, B, c.T, c.S, t_prompt, c.image_vocab_size, st));
  const int Tnew = c.T - t_prompt;
  for (int t = t_prompt; t < c.T; ++t) {
    const float* nz = noise ? noise + (int64_t)(t - t_prompt) * (steps - 1) * B * c.S : nullptr;
    const float* un = uniform ? uniform + (int64_t)(t - t_prompt) * steps * B * c.S * c.num_factored_vocabs : nullptr;
    const int from = (t == t_prompt) ? 0 : t - 1;
    GN_PROPAGATE(maskgit_impl(m, tokens, B, t, steps, unmask_mode, nz, un, m->samples_scratch, logits0, Tnew,
                              t - t_prompt, from, nullptr, nullptr, st));
  }
  return GN_OK;
}

int gn_generate_host(gn_model* m, int32_t* tokens_host, int B, int t_prompt, int steps, float temperature,
                     int unmask_mode, const float* noise_host, const float* uniform_host, void* stream) {
  GN_REQUIRE(m && tokens_host && B > 0, "gn_generate_host: invalid argument");
  // validate before any staging buffer is sized from (steps - 1)
  GN_PROPAGATE(check_generate_args(m, B, steps, temperature, unmask_mode, noise_host, uniform_host));
  GN_REQUIRE(t_prompt >= 1 && t_prompt <= m->cfg.T, "num_prompt_frames %d out of range [1, %d]", t_prompt, m->cfg.T);
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  const gn_config& c = m->cfg;
  const int64_t ntok = (int64_t)B * c.T * c.S;
  if (ntok > m->tokens_cap) {
    dev_free(m, m->tokens_dev);
    m->tokens_cap = 0;
    GN_PROPAGATE(dev_alloc(m, (void**)&m->tokens_dev, (size_t)ntok * 4));
    m->tokens_cap = ntok;
  }
  const int64_t nnoise = noise_host ? (int64_t)(c.T - t_prompt) * (steps - 1) * B * c.S : 0;
  const int64_t nuni = (uniform_host && temperature > 1e-8f)
                           ? (int64_t)(c.T - t_prompt) * steps * B * c.S * c.num_factored_vocabs : 0;
  if (nnoise + nuni > m->noise_cap) {   // one staging buffer: [noise | uniform]
    dev_free(m, m->noise_dev);
    m->noise_cap = 0;
    GN_PROPAGATE(dev_alloc(m, (void**)&m->noise_dev, (size_t)(nnoise + nuni) * 4));
    m->noise_cap = nnoise + nuni;
  }
  GN_CUDA_CHECK(cudaMemcpyAsync(m->tokens_dev, tokens_host, (size_t)ntok * 4, cudaMemcpyHostToDevice, st));
  if (nnoise) GN_CUDA_CHECK(cudaMemcpyAsync(m->noise_dev, noise_host, (size_t)nnoise * 4, cudaMemcpyHostToDevice, st));
  if (nuni)
    GN_CUDA_CHECK(cudaMemcpyAsync(m->noise_dev + nnoise, uniform_host, (size_t)nuni * 4, cudaMemcpyHostToDevice, st));
  GN_PROPAGATE(gn_generate(m, m->tokens_dev, B, t_prompt, steps, temperature, unmask_mode, nnoise ? m->noise_dev : nullptr,
                           nuni ? m->noise_dev + nnoise : nullptr, nullptr, stream));
  GN_CUDA_CHECK(cudaMemcpyAsync(tokens_host, m->tokens_dev, (size_t)ntok * 4, cudaMemcpyDeviceToHost, st));
  GN_CUDA_CHECK(cudaStreamSynchronize(st));
  return GN_OK;
}

int gn_teacher_forced_eval(gn_model* m, const int32_t* gt, int B, int steps, float temperature, int unmask_mode,
                           const float* noise, const float* uniform, int32_t* samples_out, double* acc, void* stream) {
  GN_PROPAGATE(check_generate_args(m, B, steps, temperature, unmask_mode, noise, uniform));
  if (temperature <= 1e-8f) uniform = nullptr;
  GN_REQUIRE(gt && acc, "gn_teacher_forced_eval: null buffer");
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  const gn_config& c = m->cfg;
  GN_PROPAGATE(ensure_decode(m, B));
  const int64_t TS = (int64_t)c.T * c.S;
  for (int t = 1; t < c.T; ++t) {
    // inputs_masked = GT.clone(); inputs_masked[:, t:] = mask   (evaluate.py:109-110)
    GN_CUDA_CHECK(cudaMemcpyAsync(m->prompt_scratch, gt, (size_t)B * TS * 4, cudaMemcpyDeviceToDevice, st));
    GN_PROPAGATE(fill_frames(m->prompt_scratch, B, c.T, c.S, t, c.image_vocab_size, st));
    const float* nz = noise ? noise + (int64_t)(t - 1) * (steps - 1) * B * c.S : nullptr;
    const float* un = uniform ? uniform + (int64_t)(t - 1) * steps * B * c.S * c.num_factored_vocabs : nullptr;
    int32_t* sout = m->samples_scratch;
    GN_PROPAGATE(maskgit_impl(m, m->prompt_scratch, B, t, steps, unmask_mode, nz, un, sout, nullptr, 1, 0, t - 1,
                              gt + (int64_t)t * c.S, acc, st));
    GN_PROPAGATE(launch_count_equal(sout, c.S, gt + (int64_t)t * c.S, TS, c.S, B * c.S, acc, st));
    if (samples_out) {
      GN_CUDA_CHECK(cudaMemcpy2DAsync(samples_out + (int64_t)(t - 1) * c.S, (size_t)(c.T - 1) * c.S * 4, sout,
                                      (size_t)c.S * 4, (size_t)c.S * 4, B, cudaMemcpyDeviceToDevice, st));
    }
  }
  return GN_OK;
}

int gn_forward_loss(gn_model* m, const int32_t* input_ids, const int32_t* labels, int B, float* logits, double* acc,
                    void* stream) {
  GN_REQUIRE(m && input_ids && labels && acc && B > 0, "gn_forward_loss: invalid argument");
  GN_PROPAGATE(gn_model_check_weights(m));
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  const gn_config& c = m->cfg;
  GN_PROPAGATE(ensure_decode(m, B));
  const int TS = c.T * c.S;
  {
    const int64_t total = (int64_t)B * TS;
    const int grid = (int)std::min<int64_t>(ceil_div64(total, 256), 1184);
    relevant_weight_kernel<<<grid, 256, 0, st>>>(input_ids, m->weight, total, TS, c.S, c.image_vocab_size);
    GN_CUDA_CHECK(cudaGetLastError());
    ++g_launch_count;
  }
  const int cc = chunk_clips_for(m, c.T);
  for (int b0 = 0; b0 < B; b0 += cc) {
    const int nb = std::min(cc, B - b0);
    GN_PROPAGATE(forward_chunk(m, input_ids, b0, nb, 0, c.T, false, st));
    GN_PROPAGATE(ensure_rows(m, (int64_t)nb * TS));
    GN_PROPAGATE(readout(m, nb, c.T, -1, m->rows, st));
    GN_PROPAGATE(launch_ce(m->rows, labels + (int64_t)b0 * TS, TS, TS, nb * TS, c.factored_vocab_size,
                           c.num_factored_vocabs, m->weight + (int64_t)b0 * TS, acc, st));
    if (logits)
      GN_PROPAGATE(launch_logits_transpose(m->rows, logits + (int64_t)b0 * m->C * TS, nb, c.T, c.S, m->C, c.T, 0, st));
  }
  return GN_OK;
}

int gn_spatial_attention(const void* qkv, void* out, int n_frames, int S, int n_heads, int head_dim, float scale,
                         int kernel, void* stream) {
  GN_REQUIRE(qkv && out && n_frames > 0 && S > 0 && n_heads > 0 && head_dim > 0, "gn_spatial_attention: invalid argument");
  AttnArgs aa{};
  aa.qkv = qkv; aa.out = out; aa.act_bf16 = 1; aa.n_heads = n_heads; aa.head_dim = head_dim; aa.scale = scale;
  aa.fp16 = (kernel & 0x100) ? 1 : 0;   // bit 8: the buffers hold IEEE fp16 instead of bf16
  kernel &= 0xff;
  cudaStream_t st = (cudaStream_t)stream;
  if (kernel == 1) {
    GN_REQUIRE(fast_spatial_supported(aa, S), "mma.sync spatial kernel does not support this shape");
    return fast_spatial_attention(aa, n_frames, S, st);
  }
  return launch_spatial_attention(aa, n_frames, S, kernel == 2, st);
}

int gn_linear_forward(const void* a, const void* w, const float* bias, const float* resid, void* out, void* out2, int M,
                      int N, int K, int epi, int in_bf16, int out_bf16, int force_simt, void* stream) {
  LinearArgs la{};
  la.A = a; la.lda = K; la.W = w; la.ldw = K; la.bias = bias; la.resid = resid; la.ldr = N;
  la.out = out; la.ldo = N; la.out2 = out2; la.ldo2 = N;
  la.M = M; la.N = N; la.K = K; la.epi = epi; la.in_bf16 = in_bf16 != 0; la.out_bf16 = out_bf16 != 0;
  la.fp16 = in_bf16 == 2 || out_bf16 == 2;   // 2 = IEEE fp16 instead of bf16
  la.force_simt = force_simt;
  return linear_forward(la, (cudaStream_t)stream);
}

int gn_sample_tokens(const float* logits_rows, int R, int V, int NV, const float* uniform, int32_t* samples,
                     float* conf, void* stream) {
  GN_REQUIRE(logits_rows && samples && conf && R > 0, "gn_sample_tokens: invalid argument");
  return launch_sample(logits_rows, R, V, NV, uniform, samples, conf, (cudaStream_t)stream);
}
int gn_remask_step(int32_t* prompt_frame, int64_t clip_stride, const int32_t* samples, const float* conf_or_noise,
                   uint8_t* unmasked, int32_t* samples_out, int B, int S, int n_mask, int last_step, int mask_id,
                   void* stream) {
  GN_REQUIRE(prompt_frame && samples && unmasked && samples_out && B > 0 && S > 0, "gn_remask_step: invalid argument");
  return launch_remask(prompt_frame, clip_stride, samples, conf_or_noise, unmasked, samples_out, B, S, n_mask, last_step,
                       mask_id, (cudaStream_t)stream);
}
int gn_cross_entropy(const float* logits_rows, const int32_t* targets, int R, int V, int NV, const uint8_t* weight,
                     double* acc, void* stream) {
  GN_REQUIRE(logits_rows && targets && acc && R > 0, "gn_cross_entropy: invalid argument");
  return launch_ce(logits_rows, targets, 0, R, R, V, NV, weight, acc, (cudaStream_t)stream);
}

int gn_profile_begin(void) {
  gn::g_gemm_flops_issued = 0.0;
  return gn::profile_begin();
}
int gn_profile_end(double* out) {
  GN_REQUIRE(out, "gn_profile_end: null output");
  GN_PROPAGATE(gn::profile_end(out));
  out[2 * gn::PC_COUNT] = gn::g_gemm_flops_issued;
  return GN_OK;
}

double gn_model_flops_per_clip_forward(gn_model* m) {
  if (!m) return 0.0;
  const gn_config& c = m->cfg;
  const double d = c.d_model;
  const double per_tok_layer = 32.0 * d * d + 4.0 * d * (c.S + c.T);
  return (c.num_layers * per_tok_layer + 2.0 * d * m->C) * c.T * c.S;
}
double gn_model_flops_executed(gn_model* m) { return m ? m->flops_executed : 0.0; }
double gn_model_bytes_executed(gn_model* m) { return m ? m->bytes_executed : 0.0; }
void gn_model_reset_counters(gn_model* m) {
  if (m) m->flops_executed = m->bytes_executed = 0.0;
}

}  // extern "C"
