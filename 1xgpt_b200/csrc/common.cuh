// Shared host/device helpers for libgenie_b200.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace gn {

typedef __nv_bfloat16 bf16;
typedef __half f16;

// The 16-bit tensor-core paths are written once and instantiated for both 16-bit formats (tcgen05 kind::f16 takes
// either): bf16 (8-bit mantissa, fp32 range) and IEEE fp16 (11-bit mantissa = the tf32 mantissa, +-65504).  fp16 is the
// parity format: with fp32 accumulation / residual / LayerNorm / softmax it meets the 1e-3 logits bar that bf16 cannot
// (scripts/precision_budget.py); conversions saturate to the largest finite value instead of overflowing to inf.
template <typename H> struct H16;
template <> struct H16<bf16> {
  static constexpr uint32_t UMMA_FMT = 1;   // UMMA_FMT_BF16
  static constexpr CUtensorMapDataType TMAP = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
};
template <> struct H16<f16> {
  static constexpr uint32_t UMMA_FMT = 0;   // UMMA_FMT_F16
  static constexpr CUtensorMapDataType TMAP = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
};
template <> struct H16<float> {            // fp32 operands run as kind::tf32
  static constexpr uint32_t UMMA_FMT = 2;   // UMMA_FMT_TF32
  static constexpr CUtensorMapDataType TMAP = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
};

// ---- error plumbing (thread-local message, negative return codes; never throws across the C ABI)
enum : int {
  GN_OK = 0,
  GN_ERR_INVALID = -1,
  GN_ERR_CUDA = -2,
  GN_ERR_UNSUPPORTED = -3,
  GN_ERR_STATE = -4,
};

void set_error(const char* fmt, ...);
const char* last_error();

#define GN_CUDA_CHECK(expr)                                                                     \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::gn::set_error("%s:%d CUDA error %d (%s) in `%s`", __FILE__, __LINE__, (int)_e,          \
                      cudaGetErrorString(_e), #expr);                                           \
      return ::gn::GN_ERR_CUDA;                                                                 \
    }                                                                                           \
  } while (0)

#define GN_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      ::gn::set_error(__VA_ARGS__);      \
      return ::gn::GN_ERR_INVALID;       \
    }                                    \
  } while (0)

#define GN_PROPAGATE(expr)      \
  do {                          \
    int _rc = (expr);           \
    if (_rc != 0) return _rc;   \
  } while (0)

// Programmatic dependent launch (PDL).  Kernels of the forward chain are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (when enabled): kernel N+1 may be scheduled and run its
// prologue (barrier init, TMEM allocation, descriptor prefetch) while kernel N drains; it must not touch global
// memory before pdl_wait().  pdl_wait() is a no-op for a kernel launched without the attribute.
extern bool g_use_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Live per-launch timing (bench.py): while profiling is on, every launch_kernel() call is bracketed by CUDA
// events on its stream and accounted to a category.
enum ProfCat : int { PC_GEMM_STORE = 0, PC_GEMM_GELU, PC_GEMM_RESID, PC_PREP, PC_SPATIAL, PC_TEMPORAL, PC_OTHER, PC_COUNT };
void prof_before(int cat, cudaStream_t st);
void prof_after(cudaStream_t st);
extern bool g_prof_on;

// L2 residency of the fp32 residual stream.  While a window is set (model.cu sets it around the layer stack of a
// chunk), every kernel launch carries cudaLaunchAttributeAccessPolicyWindow for that address range with the
// "persisting" property: the stream is read and rewritten by 3 residual GEMMs and read by 2 LayerNorm passes per block,
// so keeping it in the L2 set-aside removes most of its HBM traffic (the wide activations - QKV, MLP hidden - stream
// through the rest of L2).  num_bytes == 0: no attribute.
extern thread_local cudaAccessPolicyWindow g_l2_window;   // per host thread: handles on different threads do not interfere
inline int add_l2_window_attr(cudaLaunchAttribute* attr, int n) {
  if (g_l2_window.num_bytes == 0) return n;
  attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
  attr[n].val.accessPolicyWindow = g_l2_window;
  return n + 1;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(int cat, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  if (g_prof_on) prof_before(cat, st);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  na = add_l2_window_attr(attr, na);
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (g_prof_on) prof_after(st);
  return e;
}

// same, as thread-block clusters of `cluster_x` CTAs along x (grid.x must be a multiple of cluster_x)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(int cat, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t st, int cluster_x, Args&&... args) {
  if (g_prof_on) prof_before(cat, st);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = cluster_x;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (g_use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  na = add_l2_window_attr(attr, na);
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (g_prof_on) prof_after(st);
  return e;
}

// ---- per-device one-time setup.  cudaFuncSetAttribute / cudaDeviceSetLimit and the SM count belong to the device that
// is current when they are called; the C ABI takes a device index per handle, so every such cache is keyed by the
// device ordinal (a second handle on another GPU of the same process gets its own opt-ins).
constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}
int device_sm_count();   // SMs of the current device (runtime.cu)
struct DevSmemOptIn { int set[kMaxDevices] = {}; };
// raise the dynamic-shared-memory limit of `kern` on the current device to at least `bytes` (no-op once done)
template <typename K>
inline cudaError_t ensure_smem_optin(DevSmemOptIn& s, K kern, int bytes) {
  const int dev = current_device();
  if (bytes <= s.set[dev]) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) s.set[dev] = bytes;
  return e;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<f16>(f16 v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ f16 from_f32<f16>(float v) {
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
}

// two fp32 -> one packed 16-bit pair (lo in bits [0,16)), round to nearest even
template <typename H>
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack_h2<bf16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <>
__device__ __forceinline__ uint32_t pack_h2<f16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <typename H>
__device__ __forceinline__ float2 unpack_h2(uint32_t v);
template <>
__device__ __forceinline__ float2 unpack_h2<bf16>(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
template <>
__device__ __forceinline__ float2 unpack_h2<f16>(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// round-to-nearest fp32 -> tf32 (tcgen05 kind::tf32 TRUNCATES its fp32 operands, which biases every product
// by ~2^-11; operands are therefore pre-rounded wherever they are produced in the tf32 parity mode)
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// exact-erf GELU (nn.GELU() default, st_transformer.py:17)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }


// GELU for bf16 outputs: erfc(|x|/sqrt2) = 2^q(|x|) with a degree-5 polynomial q (no constant term), fitted on
// [0, 6] (weighted minimax, scripts in DESIGN.md); max |gelu_fast - gelu_erf| = 5.4e-7, relative error <= 4e-4
// wherever |gelu| > 1e-3 (and ~1e-6 for x > 0): two orders below the bf16 rounding of the stored value.
// 1 MUFU (ex2) + 9 FP32 ops instead of erff's ~40.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fminf(fabsf(x), 6.0f);
  float q = fmaf(-4.88107450e-04f, z, 7.19873621e-03f);
  q = fmaf(q, z, -5.21466308e-02f);
  q = fmaf(q, z, -4.59595885e-01f);
  q = fmaf(q, z, -1.15100052e+00f);
  q *= z;
  float e;                             // erfc(|x|/sqrt2) in (0, 1]; q in [-31, 0]: no denormal handling needed
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  const float h = 0.5f * x;
  const float ah = fabsf(h);
  return h + fmaf(-ah, e, ah);         // 0.5x + 0.5|x| (1 - erfc)
}


// ---- packed fp32x2 arithmetic (sm_100: one FFMA2 does two FMAs; the scalar 3-register FFMA issues at half rate)
__device__ __forceinline__ uint64_t pack_f2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// gelu_fast on two values at once (same polynomial), fma-pipe work done with packed instructions
__device__ __forceinline__ void gelu_fast2(float& x0, float& x1) {
  const float z0 = fminf(fabsf(x0), 6.0f), z1 = fminf(fabsf(x1), 6.0f);
  const uint64_t z = pack_f2(z0, z1);
  uint64_t q = fma2(pack_f2(-4.88107450e-04f, -4.88107450e-04f), z, pack_f2(7.19873621e-03f, 7.19873621e-03f));
  q = fma2(q, z, pack_f2(-5.21466308e-02f, -5.21466308e-02f));
  q = fma2(q, z, pack_f2(-4.59595885e-01f, -4.59595885e-01f));
  q = fma2(q, z, pack_f2(-1.15100052e+00f, -1.15100052e+00f));
  q = mul2(q, z);
  float q0, q1, e0, e1;
  unpack_f2(q, q0, q1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
  const uint64_t x = pack_f2(x0, x1);
  const uint64_t h = mul2(x, pack_f2(0.5f, 0.5f));
  float h0, h1;
  unpack_f2(h, h0, h1);
  const uint64_t ah = pack_f2(fabsf(h0), fabsf(h1));
  const uint64_t t = fma2(pack_f2(e0, e1), pack_f2(-1.0f, -1.0f), pack_f2(1.0f, 1.0f));   // 1 - erfc
  unpack_f2(fma2(ah, t, h), x0, x1);
}

}  // namespace gn
