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

}  // namespace y5
