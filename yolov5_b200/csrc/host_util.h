// Host-side helpers shared by the translation units of liby5b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace y5 {

// records a thread-local message retrievable through y5_last_error(); returns `code`
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
using EncodeIm2colFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// resolved through cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda
EncodeTiledFn driver_fn_encode_tiled();
EncodeIm2colFn driver_fn_encode_im2col();


CUtensorMapSwizzle swizzle_for_row_bytes(int row_bytes);
CUtensorMapDataType tm_dtype(int dtype);
// tiled tensor map of `rank` dims (innermost first); on failure records a message and returns Y5_E_DRIVER
int encode_tiled(CUtensorMap* map, int dtype, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                 const cuuint32_t* box, CUtensorMapSwizzle sw, const char* what);
// im2col tensor map over an NHWC view (element strides xs/ys/ns), filter kh x kw, conv stride / padding
int encode_im2col(CUtensorMap* map, int dtype, const void* base, int C, int W, int H, int N, long long xs, long long ys, long long ns,
                  int kh, int kw, int stride, int pad_h, int pad_w, uint32_t channels_per_pixel, uint32_t pixels_per_column,
                  CUtensorMapSwizzle sw);

}  // namespace y5
