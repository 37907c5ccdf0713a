// Host-side CUtensorMap construction.  cuTensorMapEncodeTiled is resolved at run time through
// cudaGetDriverEntryPoint, so the library does not link against libcuda.
#pragma once
#include "common.cuh"

namespace gn {

// 2-D row-major tensor: `inner` contiguous elements per row, `outer` rows, row pitch `ld` elements.
int make_tensor_map_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, int64_t inner,
                       int64_t outer, int64_t ld, int box_inner, int box_outer, CUtensorMapSwizzle swz);

// 3-D tensor (dims innermost first), strides in elements for dims 1 and 2.
int make_tensor_map_3d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, int64_t d0,
                       int64_t d1, int64_t d2, int64_t stride1, int64_t stride2, int box0, int box1, int box2,
                       CUtensorMapSwizzle swz);

// 4-D NHWC activation tensor [N, H, W, C] (C innermost) for implicit-GEMM convolution: box {bc, bw, bh, 1} in
// traversal elements with per-dimension traversal strides (1, sw, sh, 1); out-of-bounds coordinates read zeros,
// which implements the conv padding.
int make_tensor_map_nhwc(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, int64_t N,
                         int64_t H, int64_t W, int64_t C, int box_c, int box_w, int box_h, int stride_w, int stride_h,
                         CUtensorMapSwizzle swz);

// General tiled map: `rank` dims (innermost first), byte strides for dims 1..rank-1, box per dim.
int make_tensor_map_nd(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int rank, const int64_t* dims,
                       const int64_t* strides_bytes, const int* box, CUtensorMapSwizzle swz);

}  // namespace gn
