// Host-side helpers shared by the translation units of liby5b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

namespace y5 {

// Launch with the programmatic-dependent-launch attribute: the kernel may be scheduled while its predecessor in the
// stream drains (every such kernel starts with griddepcontrol.wait, which blocks until the predecessor has completed and
// flushed, then griddepcontrol.launch_dependents).  Y5_PDL=0 turns the attribute off.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static const bool pdl = [] { const char* e = getenv("Y5_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// records a thread-local message retrievable through y5_last_error(); returns `code`
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is per device, and one process
// may drive several GPUs (DataParallel, a model on cuda:1 while cuda:0 is current ...)
cudaError_t ensure_dyn_smem(const void* kernel, int bytes);

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
