#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "../../include/y5b200.h"
#include "host_util.h"

namespace y5 {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

static void* driver_entry(const char* name) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return fn;
}
EncodeTiledFn driver_fn_encode_tiled() {
    static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_entry("cuTensorMapEncodeTiled"));
    return fn;
}
EncodeIm2colFn driver_fn_encode_im2col() {
    static EncodeIm2colFn fn = reinterpret_cast<EncodeIm2colFn>(driver_entry("cuTensorMapEncodeIm2col"));
    return fn;
}

}  // namespace y5

extern "C" Y5_API int y5_version(void) { return 1; }
extern "C" Y5_API const char* y5_last_error(void) { return y5::g_err; }
extern "C" Y5_API int64_t y5_launch_count(void) { return y5::g_launches.load(std::memory_order_relaxed); }
