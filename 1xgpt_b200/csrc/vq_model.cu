// MAGVIT2 tokenizer (Encoder -> LFQ, LFQ^-1 -> Decoder) orchestration + C ABI.
// reference: magvit2/modules/diffusionmodules/improved_model.py:54-182 (Encoder/Decoder/ResBlock/Upsampler),
//            magvit2/modules/vqvae/lookup_free_quantize.py:181-194,241-257, magvit2/config.py:12-18,
//            visualize.py:84-116 (decode wrapper: little-endian bit order, uint8 rescale).
// NHWC activations, fp32 trunk; every 3x3 / 1x1 convolution with Cin % 64 == 0 runs on the tcgen05 GEMM kernel
// (implicit GEMM: TMA boxes over the NHWC input, taps folded into the K loop, zero padding = TMA OOB fill).
#include "kernels.cuh"
#include "vq_kernels.cuh"
#include "../../include/genie_b200.h"
#include <cstdlib>
#include <map>
#include <set>
#include <string>
#include <vector>

using namespace gn;

namespace {
struct ConvW {
  void* w = nullptr;      // [Cout, taps*Cin] in the operand format (bf16 / fp16 tensor path, fp32 exact mode)
  float* w_raw = nullptr; // original fp32 OIHW (direct kernels)
  float* b = nullptr;
  int cout = 0, cin = 0, taps = 0;
};
struct NormW { float* g = nullptr; float* b = nullptr; int c = 0; };
struct ResW {
  NormW n1, n2;
  ConvW c1, c2, nin;
  int cin = 0, cout = 0;
};
}  // namespace

struct gn_vq {
  gn_vq_config cfg;
  int device = 0;
  int nb = 0;  // number of resolution levels
  std::vector<int> ch;
  // encoder
  ConvW enc_in;
  std::vector<std::vector<ResW>> enc_down;
  std::vector<ConvW> enc_ds;
  std::vector<ResW> enc_mid;
  NormW enc_norm;
  ConvW enc_out;
  // decoder
  ConvW dec_in;
  std::vector<ResW> dec_mid;
  std::vector<std::vector<ResW>> dec_up;
  std::vector<ConvW> dec_us;
  NormW dec_norm;
  ConvW dec_out;
  std::set<std::string> have;
  std::vector<void*> owned;
  // workspace for `cap_imgs` images at full resolution
  int cap_imgs = 0, cap_H = 0, cap_W = 0;
  float *x = nullptr, *y = nullptr, *r = nullptr;
  void* a = nullptr;     // convolution operand (GroupNorm + swish output / cast trunk) in the operand format
  double* stats = nullptr;
  int o16 = 2;           // operand format: 2 = fp16 (default), 1 = bf16, 0 = fp32 on the CUDA-core kernels (exact mode)
  void* dec_out_frag = nullptr;   // decoder.conv_out weights in mma fragment order (16-bit modes, launch_out_conv_pack)
  bool out_frag_dirty = true;
  unsigned int* gn_ready = nullptr;   // per-image "statistics ready" flags of the fused GroupNorm kernel
  unsigned int gn_epoch = 0;          // launch counter: the flag value of the current launch
  bool gn_fused = false;              // GENIE_B200_GN_FUSED: one persistent GroupNorm kernel (16-bit modes)
  bool out_conv_mma = true;       // GENIE_B200_OUT_CONV_MMA=0: output conv on the CUDA-core kernel (A/B)
  int per = 32;          // images per pass through the trunk (GENIE_B200_VQ_PER; workspace = per x ~120 MB at 256x256)
  size_t esz() const { return o16 ? 2 : 4; }
};

namespace {

int valloc(gn_vq* m, void** p, size_t bytes) {
  GN_CUDA_CHECK(cudaMalloc(p, bytes ? bytes : 16));
  m->owned.push_back(*p);
  return GN_OK;
}
void vfree(gn_vq* m, void* p) {
  if (!p) return;
  for (auto& q : m->owned)
    if (q == p) { q = nullptr; break; }
  cudaFree(p);
}

int ensure_ws(gn_vq* m, int imgs, int H, int W) {
  if (imgs <= m->cap_imgs && H * W <= m->cap_H * m->cap_W) return GN_OK;
  vfree(m, m->x); vfree(m, m->y); vfree(m, m->r); vfree(m, m->a); vfree(m, m->stats);
  // largest activation per image over all levels: level i holds ch[i] (or, entering from the coarser level,
  // ch[i+1]) channels at (H>>i)x(W>>i); the upsampler conv output at level i has 4*ch[i] channels
  int64_t maxc_hw = 0;
  for (int i = 0; i < m->nb; ++i) {
    const int64_t hw = (int64_t)(H >> i) * (W >> i);
    int c = std::max(m->ch[i], m->ch[std::min(i + 1, m->nb - 1)]);
    if (i > 0) c = std::max(c, 4 * m->ch[i]);
    maxc_hw = std::max(maxc_hw, hw * c);
  }
  const size_t elems = (size_t)imgs * maxc_hw;
  GN_PROPAGATE(valloc(m, (void**)&m->x, elems * 4));
  GN_PROPAGATE(valloc(m, (void**)&m->y, elems * 4));
  GN_PROPAGATE(valloc(m, (void**)&m->r, elems * 4));
  GN_PROPAGATE(valloc(m, (void**)&m->a, elems * m->esz()));
  // final stats + last-block tickets (zeroed once; every launch leaves them at zero) + partials
  // (partials: up to 512 chunks x 32 groups x {sum, sumsq} per image for the fused kernel, 64 chunks for the two-kernel form)
  const size_t stat_bytes = (64 + (size_t)imgs * (64 + 512 * 64)) * sizeof(double);
  GN_PROPAGATE(valloc(m, (void**)&m->stats, stat_bytes));
  GN_CUDA_CHECK(cudaMemset(m->stats, 0, stat_bytes));
  if (!m->gn_ready) {
    GN_PROPAGATE(valloc(m, (void**)&m->gn_ready, 128 * sizeof(unsigned int)));
    GN_CUDA_CHECK(cudaMemset(m->gn_ready, 0, 128 * sizeof(unsigned int)));
  }
  m->cap_imgs = imgs; m->cap_H = H; m->cap_W = W;
  return GN_OK;
}

// out[pix, Cout] = conv(a_bf16 NHWC) (+bias) (+resid)
int conv(gn_vq* m, const ConvW& w, const void* a, int B, int Hi, int Wi, int stride, const float* resid, float* out,
         cudaStream_t st) {
  const int Ho = Hi / stride, Wo = Wi / stride;
  LinearArgs la{};
  ConvGeom g{B, Hi, Wi, w.cin, Ho, Wo, stride};
  la.A = a; la.lda = w.cin; la.W = w.w; la.ldw = (int64_t)w.taps * w.cin; la.bias = w.b;
  la.resid = resid; la.ldr = w.cout; la.out = out; la.ldo = w.cout; la.out2 = nullptr; la.ldo2 = w.cout;
  la.M = B * Ho * Wo; la.N = w.cout; la.K = w.taps * w.cin;
  la.epi = resid ? EPI_RESID : EPI_STORE; la.in_bf16 = m->o16 != 0; la.fp16 = m->o16 == 2; la.out_bf16 = 0;
  la.force_simt = m->o16 == 0;
  la.conv = w.taps == 9 ? &g : nullptr;
  return linear_forward(la, st);
}

// swish(GroupNorm(x)) -> m->a in the operand format: one persistent kernel in the 16-bit modes (GENIE_B200_GN_FUSED),
// statistics pass + apply pass otherwise
int gn_swish(gn_vq* m, const float* x, const NormW& nw, int B, int HW, int C, cudaStream_t st) {
  if (m->gn_fused && m->o16 != 0)
    return launch_gn_swish_fused(x, m->stats, m->gn_ready, ++m->gn_epoch, nw.g, nw.b, m->a, m->o16, B, HW, C, st);
  return launch_gn_swish(x, m->stats, nw.g, nw.b, m->a, m->o16, B, HW, C, st);
}

// ResBlock (improved_model.py:36-51): x -> x + / nin(x) + conv2(swish(GN(conv1(swish(GN(x))))));  result in m->x
int resblock(gn_vq* m, const ResW& rw, int B, int H, int W, cudaStream_t st) {
  const int HW = H * W;
  GN_PROPAGATE(gn_swish(m, m->x, rw.n1, B, HW, rw.cin, st));
  GN_PROPAGATE(conv(m, rw.c1, m->a, B, H, W, 1, nullptr, m->y, st));
  const float* resid = m->x;
  if (rw.cin != rw.cout) {
    // 1x1 shortcut on the raw input (improved_model.py:34,49)
    GN_PROPAGATE(launch_prep(m->x, m->a, m->o16, nullptr, nullptr, B * HW, rw.cin, 1.f, 1, 1, -1, st));
    GN_PROPAGATE(conv(m, rw.nin, m->a, B, H, W, 1, nullptr, m->r, st));
    resid = m->r;
  }
  GN_PROPAGATE(gn_swish(m, m->y, rw.n2, B, HW, rw.cout, st));
  GN_PROPAGATE(conv(m, rw.c2, m->a, B, H, W, 1, resid, rw.cin != rw.cout ? m->x : m->x, st));
  return GN_OK;
}

int put_f32(gn_vq* m, float** dst, const float* src, int64_t n, cudaStream_t st) {
  if (!*dst) GN_PROPAGATE(valloc(m, (void**)dst, (size_t)n * 4));
  GN_CUDA_CHECK(cudaMemcpyAsync(*dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  return GN_OK;
}

int set_conv(gn_vq* m, ConvW& c, const std::string& leaf, const float* src, const int64_t* shape, int ndim, int64_t numel,
             cudaStream_t st, bool tensor_path) {
  if (leaf == "weight") {
    GN_REQUIRE(ndim == 4, "conv weight must be 4-D");
    c.cout = (int)shape[0]; c.cin = (int)shape[1]; c.taps = (int)(shape[2] * shape[3]);
    GN_PROPAGATE(put_f32(m, &c.w_raw, src, numel, st));
    if (tensor_path) {
      if (!c.w) GN_PROPAGATE(valloc(m, (void**)&c.w, (size_t)numel * m->esz()));
      GN_PROPAGATE(launch_repack_conv_w(src, c.w, m->o16, c.cout, c.cin, c.taps, st));
    }
    return GN_OK;
  }
  if (leaf == "bias") return put_f32(m, &c.b, src, numel, st);
  set_error("unknown conv parameter %s", leaf.c_str());
  return GN_ERR_INVALID;
}
int set_norm(gn_vq* m, NormW& n, const std::string& leaf, const float* src, int64_t numel, cudaStream_t st) {
  n.c = (int)numel;
  if (leaf == "weight") return put_f32(m, &n.g, src, numel, st);
  if (leaf == "bias") return put_f32(m, &n.b, src, numel, st);
  set_error("unknown norm parameter %s", leaf.c_str());
  return GN_ERR_INVALID;
}
int set_res(gn_vq* m, ResW& r, const std::string& rest, const float* src, const int64_t* shape, int ndim, int64_t numel,
            cudaStream_t st) {
  const auto dot = rest.find('.');
  const std::string mod = rest.substr(0, dot), leaf = rest.substr(dot + 1);
  if (mod == "norm1") return set_norm(m, r.n1, leaf, src, numel, st);
  if (mod == "norm2") return set_norm(m, r.n2, leaf, src, numel, st);
  if (mod == "conv1") { int rc = set_conv(m, r.c1, leaf, src, shape, ndim, numel, st, true); r.cin = r.c1.cin; r.cout = r.c1.cout; return rc; }
  if (mod == "conv2") return set_conv(m, r.c2, leaf, src, shape, ndim, numel, st, true);
  if (mod == "nin_shortcut") return set_conv(m, r.nin, leaf, src, shape, ndim, numel, st, true);
  set_error("unknown ResBlock member %s (conv_shortcut variant is not used by VQConfig)", mod.c_str());
  return GN_ERR_INVALID;
}

struct VqGuard {
  int prev = -1;
  explicit VqGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~VqGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

int gn_vq_create(gn_vq** out, const gn_vq_config* cfg, int device) {
  GN_REQUIRE(out && cfg, "gn_vq_create: null argument");
  *out = nullptr;
  GN_REQUIRE(cfg->num_blocks >= 1 && cfg->num_blocks <= 8, "num_blocks out of range");
  GN_REQUIRE(cfg->in_channels == 3 && cfg->out_channels == 3, "MAGVIT2 path supports 3-channel images");
  GN_REQUIRE(cfg->base_channels % 128 == 0, "base_channels must be a multiple of 128 (tcgen05 conv path, GroupNorm(32))");
  GN_REQUIRE(cfg->z_channels >= 1 && cfg->z_channels <= 31, "z_channels out of range");
  int ndev = 0;
  GN_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  GN_REQUIRE(device >= 0 && device < ndev, "device %d not available", device);
  gn_vq* m = new (std::nothrow) gn_vq();
  GN_REQUIRE(m, "out of host memory");
  m->cfg = *cfg;
  m->device = device;
  GN_REQUIRE(cfg->precision == GN_PREC_BF16 || cfg->precision == GN_PREC_FP16 || cfg->precision == GN_PREC_FP32,
             "MAGVIT2 path: precision must be GN_PREC_FP16, GN_PREC_BF16 or GN_PREC_FP32 (got %d)", cfg->precision);
  m->o16 = cfg->precision == GN_PREC_FP16 ? 2 : (cfg->precision == GN_PREC_BF16 ? 1 : 0);
  m->nb = cfg->num_blocks;
  {
    // Images per pass.  The coarse levels of a pass are small GEMMs (16x16 latents: 256 rows per image), so more images
    // per pass fill the 148 SMs better there; the fine levels no longer fit L2 either way (33 MB fp32 per image at
    // 256x256x128).  Results do not depend on it (every kernel treats images independently; GroupNorm statistics are
    // per image in a fixed order).  Measured on B200, 64 frames of 256x256, fp16 operands, same box: 8 / 16 / 32 images
    // per pass = 2460 / 2860 / 3017 img/s encode, 1978 / 2165 / 2250 img/s decode -> default 32 (3.7 GB of workspace).
    const char* e = getenv("GENIE_B200_VQ_PER");
    const int v = e ? atoi(e) : 0;
    if (v >= 1 && v <= 64) m->per = v;
    const char* gf = getenv("GENIE_B200_GN_FUSED");
    m->gn_fused = gf ? (gf[0] != '0') : false;
    const char* oc = getenv("GENIE_B200_OUT_CONV_MMA");
    m->out_conv_mma = oc ? (oc[0] != '0') : true;   // measured: 1549 -> 436 us per 32 images, decode 2335 -> 2500 img/s
  }
  for (int i = 0; i < m->nb; ++i) m->ch.push_back(cfg->base_channels * cfg->ch_mult[i]);
  m->enc_down.assign(m->nb, std::vector<ResW>(cfg->num_res_blocks));
  m->enc_ds.resize(m->nb);
  m->enc_mid.resize(cfg->num_res_blocks);
  m->dec_mid.resize(cfg->num_res_blocks);
  m->dec_up.assign(m->nb, std::vector<ResW>(cfg->num_res_blocks));
  m->dec_us.resize(m->nb);
  *out = m;
  return GN_OK;
}

void gn_vq_destroy(gn_vq* m) {
  if (!m) return;
  VqGuard g(m->device);
  for (void* p : m->owned)
    if (p) cudaFree(p);
  delete m;
}

int gn_vq_set_weight(gn_vq* m, const char* key, const float* src, const int64_t* shape, int ndim, void* stream) {
  GN_REQUIRE(m && key && src && shape, "gn_vq_set_weight: null argument");
  VqGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= shape[i];
  const std::string k(key);
  int i = -1, j = -1;
  char rest[128] = "";
  int rc = GN_ERR_INVALID;
  if (k.rfind("encoder.conv_in.", 0) == 0) rc = set_conv(m, m->enc_in, k.substr(16), src, shape, ndim, numel, st, false);
  else if (sscanf(key, "encoder.down.%d.block.%d.%127s", &i, &j, rest) == 3 && i >= 0 && i < m->nb && j >= 0 &&
           j < m->cfg.num_res_blocks)
    rc = set_res(m, m->enc_down[i][j], rest, src, shape, ndim, numel, st);
  else if (sscanf(key, "encoder.down.%d.downsample.%127s", &i, rest) == 2 && i >= 0 && i < m->nb)
    rc = set_conv(m, m->enc_ds[i], rest, src, shape, ndim, numel, st, true);
  else if (sscanf(key, "encoder.mid_block.%d.%127s", &j, rest) == 2 && j >= 0 && j < m->cfg.num_res_blocks)
    rc = set_res(m, m->enc_mid[j], rest, src, shape, ndim, numel, st);
  else if (k.rfind("encoder.norm_out.", 0) == 0) rc = set_norm(m, m->enc_norm, k.substr(17), src, numel, st);
  else if (k.rfind("encoder.conv_out.", 0) == 0) rc = set_conv(m, m->enc_out, k.substr(17), src, shape, ndim, numel, st, false);
  else if (k.rfind("decoder.conv_in.", 0) == 0) rc = set_conv(m, m->dec_in, k.substr(16), src, shape, ndim, numel, st, false);
  else if (sscanf(key, "decoder.mid_block.%d.%127s", &j, rest) == 2 && j >= 0 && j < m->cfg.num_res_blocks)
    rc = set_res(m, m->dec_mid[j], rest, src, shape, ndim, numel, st);
  else if (sscanf(key, "decoder.up.%d.block.%d.%127s", &i, &j, rest) == 3 && i >= 0 && i < m->nb && j >= 0 &&
           j < m->cfg.num_res_blocks)
    rc = set_res(m, m->dec_up[i][j], rest, src, shape, ndim, numel, st);
  else if (sscanf(key, "decoder.up.%d.upsample.conv1.%127s", &i, rest) == 2 && i >= 0 && i < m->nb)
    rc = set_conv(m, m->dec_us[i], rest, src, shape, ndim, numel, st, true);
  else if (k.rfind("decoder.norm_out.", 0) == 0) rc = set_norm(m, m->dec_norm, k.substr(17), src, numel, st);
  else if (k.rfind("decoder.conv_out.", 0) == 0) {
    rc = set_conv(m, m->dec_out, k.substr(17), src, shape, ndim, numel, st, false);
    m->out_frag_dirty = true;
  }
  else set_error("unknown MAGVIT2 weight key %s", key);
  if (rc == GN_OK) m->have.insert(k);
  return rc;
}

int gn_vq_check_weights(gn_vq* m, int need_encoder, int need_decoder) {
  GN_REQUIRE(m, "null handle");
  std::vector<std::string> need;
  auto res = [&](const std::string& p, int cin, int cout) {
    for (const char* s : {"norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias", "conv1.weight", "conv2.weight"})
      need.push_back(p + s);
    if (cin != cout) need.push_back(p + "nin_shortcut.weight");
  };
  const int nb = m->nb, nr = m->cfg.num_res_blocks;
  if (need_encoder) {
    need.push_back("encoder.conv_in.weight");
    int cin = m->cfg.base_channels;
    for (int i = 0; i < nb; ++i) {
      for (int j = 0; j < nr; ++j) { res("encoder.down." + std::to_string(i) + ".block." + std::to_string(j) + ".", cin, m->ch[i]); cin = m->ch[i]; }
      if (i < nb - 1) { need.push_back("encoder.down." + std::to_string(i) + ".downsample.weight"); need.push_back("encoder.down." + std::to_string(i) + ".downsample.bias"); }
    }
    for (int j = 0; j < nr; ++j) res("encoder.mid_block." + std::to_string(j) + ".", cin, cin);
    for (const char* s : {"encoder.norm_out.weight", "encoder.norm_out.bias", "encoder.conv_out.weight", "encoder.conv_out.bias"}) need.push_back(s);
  }
  if (need_decoder) {
    for (const char* s : {"decoder.conv_in.weight", "decoder.conv_in.bias", "decoder.norm_out.weight", "decoder.norm_out.bias", "decoder.conv_out.weight", "decoder.conv_out.bias"}) need.push_back(s);
    int cin = m->ch[nb - 1];
    for (int j = 0; j < nr; ++j) res("decoder.mid_block." + std::to_string(j) + ".", cin, cin);
    for (int i = nb - 1; i >= 0; --i) {
      for (int j = 0; j < nr; ++j) { res("decoder.up." + std::to_string(i) + ".block." + std::to_string(j) + ".", cin, m->ch[i]); cin = m->ch[i]; }
      if (i > 0) { need.push_back("decoder.up." + std::to_string(i) + ".upsample.conv1.weight"); need.push_back("decoder.up." + std::to_string(i) + ".upsample.conv1.bias"); }
    }
  }
  for (const auto& k : need)
    if (!m->have.count(k)) { set_error("missing MAGVIT2 weight %s", k.c_str()); return GN_ERR_STATE; }
  return GN_OK;
}

// VQModel.encode (lfqgan.py:121-125): img [B,3,H,W] fp32 in [-1,1] -> token ids [B, (H/2^(nb-1)) * (W/2^(nb-1))]
int gn_vq_encode(gn_vq* m, const float* img, int B, int H, int W, int32_t* ids, float* z_out, void* stream) {
  GN_REQUIRE(m && img && ids && B > 0, "gn_vq_encode: invalid argument");
  GN_PROPAGATE(gn_vq_check_weights(m, 1, 0));
  const int nb = m->nb, down = 1 << (nb - 1);
  GN_REQUIRE(H % down == 0 && W % down == 0, "image %dx%d must be a multiple of %d", H, W, down);
  VqGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int per = m->per;  // images per pass (workspace bound)
  for (int b0 = 0; b0 < B; b0 += per) {
    const int n = std::min(per, B - b0);
    GN_PROPAGATE(ensure_ws(m, std::min(per, B), H, W));
    GN_PROPAGATE(launch_stem_conv(img + (int64_t)b0 * 3 * H * W, m->enc_in.w_raw, m->x, n, H, W, m->cfg.base_channels, st));
    int h = H, w = W;
    for (int i = 0; i < nb; ++i) {
      for (int j = 0; j < m->cfg.num_res_blocks; ++j) GN_PROPAGATE(resblock(m, m->enc_down[i][j], n, h, w, st));
      if (i < nb - 1) {
        // downsample: 3x3 stride 2 padding 1 with bias, on the raw trunk (no norm) (improved_model.py:90,113)
        GN_PROPAGATE(launch_prep(m->x, m->a, m->o16, nullptr, nullptr, n * h * w, m->ch[i], 1.f, 1, 1, -1, st));
        GN_PROPAGATE(conv(m, m->enc_ds[i], m->a, n, h, w, 2, nullptr, m->y, st));
        std::swap(m->x, m->y);
        h /= 2; w /= 2;
      }
    }
    for (int j = 0; j < m->cfg.num_res_blocks; ++j) GN_PROPAGATE(resblock(m, m->enc_mid[j], n, h, w, st));
    GN_PROPAGATE(launch_vq_head(m->x, m->stats, m->enc_norm.g, m->enc_norm.b, m->enc_out.w_raw, m->enc_out.b,
                                ids + (int64_t)b0 * h * w, z_out ? z_out + (int64_t)b0 * m->cfg.z_channels * h * w : nullptr,
                                n, h * w, m->ch[nb - 1], m->cfg.z_channels, st));
  }
  return GN_OK;
}

// decode_latents (visualize.py:104-120): ids [B, h*w] -> image [B,3,H,W] (fp32 in ~[-1,1] and/or uint8)
int gn_vq_decode(gn_vq* m, const int32_t* ids, int B, int h0, int w0, int little_endian, float* img_f32, uint8_t* img_u8,
                 void* stream) {
  GN_REQUIRE(m && ids && B > 0 && (img_f32 || img_u8), "gn_vq_decode: invalid argument");
  GN_PROPAGATE(gn_vq_check_weights(m, 0, 1));
  GN_REQUIRE((h0 * w0) % 128 == 0 && 128 % w0 == 0 && h0 % (128 / w0) == 0, "latent %dx%d not tileable", h0, w0);
  VqGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = m->nb, up = 1 << (nb - 1);
  const int H = h0 * up, W = w0 * up;
  const int per = m->per;
  for (int b0 = 0; b0 < B; b0 += per) {
    const int n = std::min(per, B - b0);
    GN_PROPAGATE(ensure_ws(m, std::min(per, B), H, W));
    int h = h0, w = w0;
    GN_PROPAGATE(launch_vq_tail(ids + (int64_t)b0 * h * w, m->dec_in.w_raw, m->dec_in.b, m->x, n, h, w, m->cfg.z_channels,
                                m->ch[nb - 1], little_endian, st));
    for (int j = 0; j < m->cfg.num_res_blocks; ++j) GN_PROPAGATE(resblock(m, m->dec_mid[j], n, h, w, st));
    for (int i = nb - 1; i >= 0; --i) {
      for (int j = 0; j < m->cfg.num_res_blocks; ++j) GN_PROPAGATE(resblock(m, m->dec_up[i][j], n, h, w, st));
      if (i > 0) {
        // Upsampler: conv3x3 C -> 4C (+bias) then depth-to-space (improved_model.py:222-237)
        GN_PROPAGATE(launch_prep(m->x, m->a, m->o16, nullptr, nullptr, n * h * w, m->ch[i], 1.f, 1, 1, -1, st));
        GN_PROPAGATE(conv(m, m->dec_us[i], m->a, n, h, w, 1, nullptr, m->y, st));
        GN_PROPAGATE(launch_depth_to_space(m->y, m->x, n, h, w, m->ch[i], st));
        h *= 2; w *= 2;
      }
    }
    GN_PROPAGATE(gn_swish(m, m->x, m->dec_norm, n, h * w, m->ch[0], st));
    float* of = img_f32 ? img_f32 + (int64_t)b0 * 3 * H * W : nullptr;
    uint8_t* ou = img_u8 ? img_u8 + (int64_t)b0 * 3 * H * W : nullptr;
    if (m->out_conv_mma && m->o16 != 0 && m->ch[0] % 64 == 0 && w % 16 == 0) {
      // 16-bit modes: implicit GEMM on mma.sync with the weights in the operand format, like every other convolution
      if (m->out_frag_dirty) {
        if (!m->dec_out_frag) GN_PROPAGATE(valloc(m, &m->dec_out_frag, (size_t)9 * (m->ch[0] / 16) * 32 * 8));
        GN_PROPAGATE(launch_out_conv_pack(m->dec_out.w_raw, m->dec_out_frag, m->o16, m->ch[0], st));
        m->out_frag_dirty = false;
      }
      GN_PROPAGATE(launch_out_conv_mma(m->a, m->o16, m->dec_out_frag, m->dec_out.b, of, ou, n, h, w, m->ch[0], st));
    } else {
      GN_PROPAGATE(launch_out_conv(m->a, m->o16, m->dec_out.w_raw, m->dec_out.b, of, ou, n, h, w, m->ch[0], st));
    }
  }
  return GN_OK;
}

}  // extern "C"
