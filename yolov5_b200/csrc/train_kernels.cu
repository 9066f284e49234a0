// Training-mode pieces of Conv = SiLU(BN(conv(x)))  (reference models/common.py:86-88, BN eps 1e-3 / momentum 0.03 set by
// initialize_weights, models/yolo.py:259) that are not GEMMs: batch statistics, normalise + activate, and their backward.
// All of them are HBM-bound passes over a [rows = B*H*W][channels] fp16/bf16 view (NHWC, possibly a channel slice of a
// wider buffer): one thread owns 8 consecutive channels (one 16-byte access per row) and walks rows, so per-channel
// partial sums live in registers; blocks combine through shared memory and publish with fp64 atomics (the fp64 total
// keeps E[y^2] - E[y]^2 well conditioned over millions of rows).
//
// Numerics follow the reference under torch.autocast: BN math in fp32 on the low-precision conv output, its result
// rounded to the activation dtype, SiLU on that rounded value, gradients rounded to the activation dtype between ops.
#include <algorithm>
#include <cstdio>

#include "../../include/y5b200.h"
#include "common.cuh"
#include "host_util.h"

namespace y5 {

constexpr int kRedThreads = 256;
constexpr int kRowsPerBlock = 1024;

struct RowGeom {
    int cgx;   // channel groups (of 8) handled side by side by one block
    int rows;  // thread rows per block
};
static inline RowGeom row_geom(int channels) {
    const int cg = channels / 8;
    RowGeom g;
    g.cgx = cg >= 32 ? 32 : (cg >= 16 ? 16 : (cg >= 8 ? 8 : (cg >= 4 ? 4 : (cg >= 2 ? 2 : 1))));
    g.rows = kRedThreads / g.cgx;
    return g;
}

__device__ __forceinline__ void load8(const void* base, long long elem_off, bool bf16, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + elem_off);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = unpack2(w[i], bf16);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void store8(void* base, long long elem_off, bool bf16, const float (&v)[8]) {
    uint4 u;
    u.x = pack2(v[0], v[1], bf16);
    u.y = pack2(v[2], v[3], bf16);
    u.z = pack2(v[4], v[5], bf16);
    u.w = pack2(v[6], v[7], bf16);
    *reinterpret_cast<uint4*>(static_cast<uint16_t*>(base) + elem_off) = u;
}
__device__ __forceinline__ float round_lowp(float x, bool bf16) { return unpack1(pack1(x, bf16), bf16); }

// block-level combine of NV per-thread vectors of 8 channels over the thread rows, then fp64 atomics
template <int NV>
__device__ __forceinline__ void block_publish(float (&acc)[NV][8], int cgx, int nrows, int channels, double* const (&dst)[NV]) {
    __shared__ float red[kRedThreads * 8];
    const int tx = threadIdx.x % cgx, ty = threadIdx.x / cgx;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) red[(ty * cgx + tx) * 8 + i] = acc[v][i];
        __syncthreads();
        // thread (tx, ty < 8) sums channel ty of group tx over all rows
        if (ty < 8) {
            float s = 0.f;
            for (int r = 0; r < nrows; ++r) s += red[(r * cgx + tx) * 8 + ty];
            const int c = (blockIdx.x * cgx + tx) * 8 + ty;
            if (c < channels) atomicAdd(dst[v] + c, static_cast<double>(s));
        }
    }
}

// mode 0: sum, sum of squares;  mode 1: sum only
template <int MODE>
__global__ void __launch_bounds__(kRedThreads) col_stats_kernel(const void* __restrict__ y, int pitch, long long rows, int channels, int bf16,
                                                                int cgx, double* __restrict__ ws) {
    const int nrows = kRedThreads / cgx;
    const int tx = threadIdx.x % cgx, ty = threadIdx.x / cgx;
    const int cg = blockIdx.x * cgx + tx;
    const bool active = cg * 8 < channels;
    float acc[MODE == 0 ? 2 : 1][8] = {};
    const long long r0 = static_cast<long long>(blockIdx.y) * kRowsPerBlock;
    const long long r1 = min(rows, r0 + kRowsPerBlock);
    if (active)
        for (long long r = r0 + ty; r < r1; r += nrows) {
            float v[8];
            load8(y, r * pitch + cg * 8, bf16 != 0, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[0][i] += v[i];
                if (MODE == 0) acc[1][i] = fmaf(v[i], v[i], acc[1][i]);
            }
        }
    if constexpr (MODE == 0) {
        double* const dst[2] = {ws, ws + channels};
        block_publish<2>(acc, cgx, nrows, channels, dst);
    } else {
        double* const dst[1] = {ws};
        block_publish<1>(acc, cgx, nrows, channels, dst);
    }
}

__global__ void bn_finalize_kernel(const double* __restrict__ ws, long long rows, int channels, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels) return;
    const double m = ws[c] / static_cast<double>(rows);
    double var = ws[channels + c] / static_cast<double>(rows) - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = static_cast<float>(m);
    invstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(m);
    if (running_var) {
        const double unbiased = rows > 1 ? var * static_cast<double>(rows) / static_cast<double>(rows - 1) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
    }
}

__global__ void col_sum_finalize_kernel(const double* __restrict__ ws, int channels, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < channels) out[c] = static_cast<float>(ws[c]);
}

__global__ void __launch_bounds__(kRedThreads) bn_act_fwd_kernel(const void* __restrict__ y, int y_pitch, void* __restrict__ z, int z_pitch,
                                                                 long long rows, int channels, int bf16, int act, int cgx,
                                                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta) {
    const int nrows = kRedThreads / cgx;
    const int tx = threadIdx.x % cgx, ty = threadIdx.x / cgx;
    const int cg = blockIdx.x * cgx + tx;
    if (cg * 8 >= channels) return;
    float mu[8], sc[8], be[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cg * 8 + i;
        mu[i] = mean[c];
        sc[i] = invstd[c] * gamma[c];
        be[i] = beta[c];
    }
    const bool b = bf16 != 0;
    const long long r0 = static_cast<long long>(blockIdx.y) * kRowsPerBlock;
    const long long r1 = min(rows, r0 + kRowsPerBlock);
    for (long long r = r0 + ty; r < r1; r += nrows) {
        float v[8];
        load8(y, r * y_pitch + cg * 8, b, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float u = round_lowp(fmaf(v[i] - mu[i], sc[i], be[i]), b);
            v[i] = act ? u / (1.0f + expf(-u)) : u;
        }
        store8(z, r * z_pitch + cg * 8, b, v);
    }
}

// du = dz * silu'(u) with u recomputed from y (rounded like the forward), rounded to the activation dtype
__device__ __forceinline__ float act_bwd(float dz, float u, int act, bool bf16) {
    if (!act) return dz;
    const float sg = 1.0f / (1.0f + expf(-u));
    return round_lowp(dz * sg * (1.0f + u * (1.0f - sg)), bf16);
}

__global__ void __launch_bounds__(kRedThreads) bn_act_bwd_reduce_kernel(const void* __restrict__ y, int y_pitch, const void* __restrict__ dz,
                                                                        int dz_pitch, long long rows, int channels, int bf16, int act,
                                                                        int cgx, const float* __restrict__ mean,
                                                                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta, double* __restrict__ ws) {
    const int nrows = kRedThreads / cgx;
    const int tx = threadIdx.x % cgx, ty = threadIdx.x / cgx;
    const int cg = blockIdx.x * cgx + tx;
    const bool active = cg * 8 < channels;
    const bool b = bf16 != 0;
    float acc[2][8] = {};
    if (active) {
        float mu[8], is[8], ga[8], be[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = cg * 8 + i;
            mu[i] = mean[c];
            is[i] = invstd[c];
            ga[i] = gamma[c];
            be[i] = beta[c];
        }
        const long long r0 = static_cast<long long>(blockIdx.y) * kRowsPerBlock;
        const long long r1 = min(rows, r0 + kRowsPerBlock);
        for (long long r = r0 + ty; r < r1; r += nrows) {
            float v[8], g[8];
            load8(y, r * y_pitch + cg * 8, b, v);
            load8(dz, r * dz_pitch + cg * 8, b, g);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float xh = (v[i] - mu[i]) * is[i];
                const float u = round_lowp(fmaf(v[i] - mu[i], is[i] * ga[i], be[i]), b);
                const float du = act_bwd(g[i], u, act, b);
                acc[0][i] += du;
                acc[1][i] = fmaf(du, xh, acc[1][i]);
            }
        }
    }
    double* const dst[2] = {ws, ws + channels};
    block_publish<2>(acc, cgx, nrows, channels, dst);
}

__global__ void __launch_bounds__(kRedThreads) bn_act_bwd_apply_kernel(const void* __restrict__ y, int y_pitch, const void* __restrict__ dz,
                                                                       int dz_pitch, void* __restrict__ dy, int dy_pitch, long long rows,
                                                                       int channels, int bf16, int act, int cgx,
                                                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                       const double* __restrict__ ws, float* __restrict__ dgamma,
                                                                       float* __restrict__ dbeta) {
    const int nrows = kRedThreads / cgx;
    const int tx = threadIdx.x % cgx, ty = threadIdx.x / cgx;
    const int cg = blockIdx.x * cgx + tx;
    if (cg * 8 >= channels) return;
    const bool b = bf16 != 0;
    float mu[8], is[8], ga[8], be[8], db[8], dg[8];
    const float inv_rows = 1.0f / static_cast<float>(rows);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cg * 8 + i;
        mu[i] = mean[c];
        is[i] = invstd[c];
        ga[i] = gamma[c];
        be[i] = beta[c];
        db[i] = static_cast<float>(ws[c]);
        dg[i] = static_cast<float>(ws[channels + c]);
        if (blockIdx.y == 0 && ty == 0) {
            dbeta[c] = db[i];
            dgamma[c] = dg[i];
        }
        db[i] *= inv_rows;
        dg[i] *= inv_rows;
    }
    const long long r0 = static_cast<long long>(blockIdx.y) * kRowsPerBlock;
    const long long r1 = min(rows, r0 + kRowsPerBlock);
    for (long long r = r0 + ty; r < r1; r += nrows) {
        float v[8], g[8];
        load8(y, r * y_pitch + cg * 8, b, v);
        load8(dz, r * dz_pitch + cg * 8, b, g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xh = (v[i] - mu[i]) * is[i];
            const float u = round_lowp(fmaf(v[i] - mu[i], is[i] * ga[i], be[i]), b);
            const float du = act_bwd(g[i], u, act, b);
            v[i] = (du - db[i] - xh * dg[i]) * ga[i] * is[i];
        }
        store8(dy, r * dy_pitch + cg * 8, b, v);
    }
}

// out[n, 2y, 2x, :] = in[n, y, x, :], every other pixel zero: turns the data gradient of a stride-2 conv into a
// stride-1 conv over the stuffed tensor
__global__ void zero_stuff2x_kernel(const uint4* __restrict__ in, int in_pitch16, uint4* __restrict__ out, int out_pitch16, int batch, int h,
                                    int w, int c16) {
    const long long total = static_cast<long long>(batch) * (2 * h) * (2 * w) * c16;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int cc = static_cast<int>(i % c16);
        long long px = i / c16;
        const int ox = static_cast<int>(px % (2 * w));
        px /= 2 * w;
        const int oy = static_cast<int>(px % (2 * h));
        const int n = static_cast<int>(px / (2 * h));
        uint4 v = make_uint4(0, 0, 0, 0);
        if (!(ox & 1) && !(oy & 1)) v = in[((static_cast<long long>(n) * h + (oy >> 1)) * w + (ox >> 1)) * in_pitch16 + cc];
        out[((static_cast<long long>(n) * 2 * h + oy) * 2 * w + ox) * out_pitch16 + cc] = v;
    }
}

static int check_view(const void* p, int pitch, int channels, const char* what) {
    if (!p) return set_error(Y5_E_INVALID, "%s: null pointer", what);
    if ((reinterpret_cast<uintptr_t>(p) & 15) || (pitch % 8) || (channels % 8) || channels <= 0 || pitch < channels)
        return set_error(Y5_E_INVALID, "%s: views must be 16-byte aligned, channels and pitch multiples of 8", what);
    return 0;
}
static dim3 row_grid(const RowGeom& g, int channels, long long rows) {
    return dim3((channels / 8 + g.cgx - 1) / g.cgx, static_cast<unsigned>((rows + kRowsPerBlock - 1) / kRowsPerBlock));
}
static int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(int(e), "%s launch failed: %s", what, cudaGetErrorString(e));
    return 0;
}

}  // namespace y5

using namespace y5;

extern "C" Y5_API int64_t y5_bn_workspace_bytes(int32_t channels) { return static_cast<int64_t>(channels) * 2 * sizeof(double); }

extern "C" Y5_API int y5_bn_stats(const void* y, int32_t pitch, int64_t rows, int32_t channels, int32_t dtype, float eps, float momentum,
                                  float* mean, float* invstd, float* running_mean, float* running_var, void* workspace, void* stream) {
    if (int e = check_view(y, pitch, channels, "bn_stats")) return e;
    if (dtype != Y5_F16 && dtype != Y5_BF16) return set_error(Y5_E_UNSUPPORTED, "bn_stats: dtype must be fp16 or bf16");
    if (!mean || !invstd || !workspace || rows <= 0) return set_error(Y5_E_INVALID, "bn_stats: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(workspace, 0, y5_bn_workspace_bytes(channels), st);
    const RowGeom g = row_geom(channels);
    count_launch(2);
    col_stats_kernel<0><<<row_grid(g, channels, rows), kRedThreads, 0, st>>>(y, pitch, rows, channels, dtype == Y5_BF16, g.cgx,
                                                                              static_cast<double*>(workspace));
    bn_finalize_kernel<<<(channels + 127) / 128, 128, 0, st>>>(static_cast<const double*>(workspace), rows, channels, eps, momentum, mean,
                                                                invstd, running_mean, running_var);
    return launch_status("bn_stats");
}

extern "C" Y5_API int y5_col_sum(const void* y, int32_t pitch, int64_t rows, int32_t channels, int32_t dtype, float* out, void* workspace,
                                 void* stream) {
    if (int e = check_view(y, pitch, channels, "col_sum")) return e;
    if (dtype != Y5_F16 && dtype != Y5_BF16) return set_error(Y5_E_UNSUPPORTED, "col_sum: dtype must be fp16 or bf16");
    if (!out || !workspace || rows <= 0) return set_error(Y5_E_INVALID, "col_sum: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(workspace, 0, static_cast<size_t>(channels) * sizeof(double), st);
    const RowGeom g = row_geom(channels);
    count_launch(2);
    col_stats_kernel<1><<<row_grid(g, channels, rows), kRedThreads, 0, st>>>(y, pitch, rows, channels, dtype == Y5_BF16, g.cgx,
                                                                              static_cast<double*>(workspace));
    col_sum_finalize_kernel<<<(channels + 127) / 128, 128, 0, st>>>(static_cast<const double*>(workspace), channels, out);
    return launch_status("col_sum");
}

extern "C" Y5_API int y5_bn_act_fwd(const void* y, int32_t y_pitch, void* z, int32_t z_pitch, int64_t rows, int32_t channels, int32_t dtype,
                                    const float* mean, const float* invstd, const float* gamma, const float* beta, int32_t act,
                                    void* stream) {
    if (int e = check_view(y, y_pitch, channels, "bn_act_fwd y")) return e;
    if (int e = check_view(z, z_pitch, channels, "bn_act_fwd z")) return e;
    if (dtype != Y5_F16 && dtype != Y5_BF16) return set_error(Y5_E_UNSUPPORTED, "bn_act_fwd: dtype must be fp16 or bf16");
    if (!mean || !invstd || !gamma || !beta || rows <= 0) return set_error(Y5_E_INVALID, "bn_act_fwd: bad argument");
    const RowGeom g = row_geom(channels);
    count_launch();
    bn_act_fwd_kernel<<<row_grid(g, channels, rows), kRedThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        y, y_pitch, z, z_pitch, rows, channels, dtype == Y5_BF16, act, g.cgx, mean, invstd, gamma, beta);
    return launch_status("bn_act_fwd");
}

extern "C" Y5_API int y5_bn_act_bwd(const void* y, int32_t y_pitch, const void* dz, int32_t dz_pitch, void* dy, int32_t dy_pitch, int64_t rows,
                                    int32_t channels, int32_t dtype, const float* mean, const float* invstd, const float* gamma,
                                    const float* beta, int32_t act, float* dgamma, float* dbeta, void* workspace, void* stream) {
    if (int e = check_view(y, y_pitch, channels, "bn_act_bwd y")) return e;
    if (int e = check_view(dz, dz_pitch, channels, "bn_act_bwd dz")) return e;
    if (int e = check_view(dy, dy_pitch, channels, "bn_act_bwd dy")) return e;
    if (dtype != Y5_F16 && dtype != Y5_BF16) return set_error(Y5_E_UNSUPPORTED, "bn_act_bwd: dtype must be fp16 or bf16");
    if (!mean || !invstd || !gamma || !beta || !dgamma || !dbeta || !workspace || rows <= 0)
        return set_error(Y5_E_INVALID, "bn_act_bwd: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(workspace, 0, y5_bn_workspace_bytes(channels), st);
    const RowGeom g = row_geom(channels);
    const dim3 grid = row_grid(g, channels, rows);
    count_launch(2);
    bn_act_bwd_reduce_kernel<<<grid, kRedThreads, 0, st>>>(y, y_pitch, dz, dz_pitch, rows, channels, dtype == Y5_BF16, act, g.cgx, mean, invstd,
                                                           gamma, beta, static_cast<double*>(workspace));
    bn_act_bwd_apply_kernel<<<grid, kRedThreads, 0, st>>>(y, y_pitch, dz, dz_pitch, dy, dy_pitch, rows, channels, dtype == Y5_BF16, act, g.cgx,
                                                          mean, invstd, gamma, beta, static_cast<const double*>(workspace), dgamma, dbeta);
    return launch_status("bn_act_bwd");
}

extern "C" Y5_API int y5_zero_stuff2x(const void* x, int32_t x_pitch, void* y, int32_t y_pitch, int32_t batch, int32_t h, int32_t w, int32_t c,
                                      int32_t dtype, void* stream) {
    if (int e = check_view(x, x_pitch, c, "zero_stuff2x x")) return e;
    if (int e = check_view(y, y_pitch, c, "zero_stuff2x y")) return e;
    if (dtype != Y5_F16 && dtype != Y5_BF16) return set_error(Y5_E_UNSUPPORTED, "zero_stuff2x: dtype must be fp16 or bf16");
    if (batch <= 0 || h <= 0 || w <= 0) return set_error(Y5_E_INVALID, "zero_stuff2x: bad shape");
    const long long total = static_cast<long long>(batch) * 4 * h * w * (c / 8);
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
    count_launch();
    zero_stuff2x_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(x), x_pitch / 8, static_cast<uint4*>(y),
                                                                               y_pitch / 8, batch, h, w, c / 8);
    return launch_status("zero_stuff2x");
}
