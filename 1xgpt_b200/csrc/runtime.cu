// Error plumbing + tensor-map encoding (host only).
#include "common.cuh"
#include "tensormap.cuh"
#include <cudaTypedefs.h>
#include <mutex>
#include <cstdlib>
#include <vector>

namespace gn {

static thread_local char g_err[1024] = "";

static bool env_flag(const char* name, bool dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  return !(v[0] == '0' || v[0] == 'n' || v[0] == 'N' || v[0] == 'f' || v[0] == 'F');
}
bool g_use_pdl = env_flag("GENIE_B200_PDL", true);

// ---- live launch profiler
bool g_prof_on = false;
thread_local cudaAccessPolicyWindow g_l2_window = {};
namespace {
struct Prof {
  std::vector<cudaEvent_t> ev;
  std::vector<int> cat;
  size_t used = 0;
} g_prof;
}  // namespace
void prof_before(int cat, cudaStream_t st) {
  while (g_prof.ev.size() < g_prof.used + 2) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    g_prof.ev.push_back(e);
  }
  if (g_prof.cat.size() < g_prof.used / 2 + 1) g_prof.cat.resize(g_prof.used / 2 + 1);
  g_prof.cat[g_prof.used / 2] = cat;
  cudaEventRecord(g_prof.ev[g_prof.used], st);
}
void prof_after(cudaStream_t st) {
  cudaEventRecord(g_prof.ev[g_prof.used + 1], st);
  g_prof.used += 2;
}
int profile_begin() {
  g_prof.used = 0;
  g_prof_on = true;
  return GN_OK;
}
// out[2*c] = total ms of category c, out[2*c+1] = launches of category c   (c < PC_COUNT)
int profile_end(double* out) {
  g_prof_on = false;
  for (int c = 0; c < 2 * PC_COUNT; ++c) out[c] = 0.0;
  for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
    GN_CUDA_CHECK(cudaEventSynchronize(g_prof.ev[i + 1]));
    float t = 0.f;
    GN_CUDA_CHECK(cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]));
    const int c = g_prof.cat[i / 2];
    out[2 * c] += t;
    out[2 * c + 1] += 1.0;
  }
  return GN_OK;
}

int device_sm_count() {
  static int sms[kMaxDevices] = {};
  const int dev = current_device();
  if (!sms[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = n > 0 ? n : 148;
    // experiment: GENIE_B200_SM_DIV=k caps every persistent grid at SMs / k, so that `lanes = k` concurrent clip streams
    // run side by side on disjoint SM sets (one lane's kernel drain / fill overlaps the other lane's compute)
    const char* e = getenv("GENIE_B200_SM_DIV");
    const int div = e ? atoi(e) : 1;
    if (div > 1) sms[dev] = (sms[dev] / div) & ~1;
  }
  return sms[dev];
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int rank, const cuuint64_t* dims,
                  const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapSwizzle swz,
                  const cuuint32_t* elem_strides = nullptr) {
  EncodeTiledFn fn = get_encode();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return GN_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = fn(out, dt, rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu] box [%u,%u,%u] base %p",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, base);
    return GN_ERR_CUDA;
  }
  return GN_OK;
}

int make_tensor_map_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, int64_t inner,
                       int64_t outer, int64_t ld, int box_inner, int box_outer, CUtensorMapSwizzle swz) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  return encode(out, base, dt, 2, dims, strides, box, swz);
}

int make_tensor_map_3d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, int64_t d0,
                       int64_t d1, int64_t d2, int64_t stride1, int64_t stride2, int box0, int box1, int box2,
                       CUtensorMapSwizzle swz) {
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)stride1 * elem_bytes, (cuuint64_t)stride2 * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, (cuuint32_t)box2};
  return encode(out, base, dt, 3, dims, strides, box, swz);
}

int make_tensor_map_nd(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int rank, const int64_t* dims,
                       const int64_t* strides_bytes, const int* box, CUtensorMapSwizzle swz) {
  cuuint64_t gd[5] = {1, 1, 1, 1, 1}, gs[4] = {0, 0, 0, 0};
  cuuint32_t bx[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gd[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
    if (i > 0) gs[i - 1] = (cuuint64_t)strides_bytes[i - 1];
  }
  return encode(out, base, dt, rank, gd, gs, bx, swz);
}

int make_tensor_map_nhwc(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, int64_t N,
                         int64_t H, int64_t W, int64_t C, int box_c, int box_w, int box_h, int stride_w, int stride_h,
                         CUtensorMapSwizzle swz) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * elem_bytes, (cuuint64_t)W * C * elem_bytes, (cuuint64_t)H * W * C * elem_bytes};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * stride_w), (cuuint32_t)(box_h * stride_h), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1};
  return encode(out, base, dt, 4, dims, strides, box, swz, es);
}

}  // namespace gn
