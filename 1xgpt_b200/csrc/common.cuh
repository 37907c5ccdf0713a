// Shared host/device helpers for libgenie_b200.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

namespace gn {

typedef __nv_bfloat16 bf16;

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

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// round-to-nearest fp32 -> tf32 (tcgen05 kind::tf32 TRUNCATES its fp32 operands, which biases every product
// by ~2^-11; operands are therefore pre-rounded wherever they are produced in the tf32 parity mode)
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// exact-erf GELU (nn.GELU() default, st_transformer.py:17)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

}  // namespace gn
