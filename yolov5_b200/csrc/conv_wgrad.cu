// Weight gradient of a convolution for sm_100a:
//     dW[co][r][s][ci] = sum over output pixels m of  dY[m][co] * X[pixel(m) shifted by tap (r,s)][ci]
// i.e. per filter tap a GEMM  dW_tap[Cout, Cin] = dY[M, Cout]^T * X_tap[M, Cin]  whose reduction dimension is the pixel
// index -- the dimension that is OUTERMOST in the NHWC activations.  Both operands are therefore fed to tcgen05.mma as
// MN-major shared-memory tiles: a TMA box of [64 pixels][64 channels] (128-byte rows, 128-byte swizzle) is exactly the
// canonical MN-major SW128 layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- each pixel is one 128-byte line
// of 64 channels, 8 pixels form a 1024-byte swizzle atom (SBO), further 64-channel blocks sit LBO bytes apart.  So
// neither dY nor X is ever transposed in memory; the tensor core does it on the fly.
//   A (dY)   : 1-2 boxes [P px][64 co] per stage -> UMMA M = 128 output channels (the upper box is not fetched when the
//              layer has <= 64 of them left: its accumulator rows are never read)
//   B (X_tap): n boxes [P px][64 ci] per stage per tap -> UMMA N = 64*n input channels (n <= 4); hardware im2col for
//              k > 1 or strided convs (same tensor map type as the forward kernel), plain 2-D tiles for 1x1/s1
// One CTA = one (co tile, tap group, ci tile, pixel range) work item.  A tap group is up to G filter taps whose
// accumulators sit side by side in TMEM (G * 64n <= 512 columns): the dY tile of a pixel block is fetched once and
// multiplied with the G shifted X tiles, which divides the dY traffic of 3x3 layers by G.  P (64/128/256 pixels per
// pipeline stage) grows when channels are few so that a stage stays ~32-64 KB and the per-stage barrier / TMA issue
// costs are amortised.  The [128 x 64n] fp32 tiles are added to the fp32 gradient with vector reductions
// (red.global.add.v4.f32); the pixel range is split so the grid is about one CTA per SM (one wave).
// Summation order across pixel ranges is not fixed (fp32 atomics), like cuDNN's default wgrad.
//
// Gradient of reference models/common.py:86-88 (Conv.forward, the nn.Conv2d weight) / models/yolo.py:97 (Detect.m[i]).
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "../../include/y5b200.h"
#include "common.cuh"
#include "host_util.h"

namespace y5 {

constexpr int kWgThreads = 64 + 128;  // warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue
constexpr int kWgCo = 128;            // output channels per tile (UMMA M)
constexpr int kWgStagesMax = 6;

struct WgradParams {
    int M, Cout, Cin;
    int kh, kw, stride, pad_h, pad_w, Wo, HoWo;
    int linear;                 // 1x1 / stride 1 / no padding: X tiles are plain 2-D boxes of the [M, Cin] matrix
    int n_blocks;               // 64-channel blocks of Cin per tile
    int ci_tiles, taps;
    int group, tap_groups;      // taps per CTA, number of tap groups
    int pix;                    // pixels (GEMM K) per pipeline stage: 64 | 128 | 256
    int kblocks, splits, kb_per_split;
    int stages;
    uint32_t box_bytes, stage_bytes;
    uint32_t idesc, tmem_cols;
    float* dw;
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
    uint64_t* empty = full + kWgStagesMax;
    uint64_t* acc_full = empty + kWgStagesMax;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDy);
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < kWgStagesMax; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    griddep_wait();  // PDL: barrier init / TMEM allocation above overlap the predecessor's tail
    griddep_launch_dependents();

    // work item
    int t = blockIdx.x;
    const int split = t % p.splits;
    t /= p.splits;
    const int ci_tile = t % p.ci_tiles;
    t /= p.ci_tiles;
    const int tg = t % p.tap_groups;
    const int co_tile = t / p.tap_groups;
    const int tap0 = tg * p.group;
    const int ntaps = min(p.group, p.taps - tap0);
    const int co0 = co_tile * kWgCo;
    const int a_boxes = p.Cout - co0 > 64 ? 2 : 1;
    const int bn = p.n_blocks * 64;
    const int ci0 = ci_tile * bn;
    const int kb0 = split * p.kb_per_split;
    const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
    const int nkb = kb1 - kb0;
    // stage layout: [A box 0][A box 1][tap 0: n boxes][tap 1: n boxes]...
    const uint32_t tx_bytes = (a_boxes + ntaps * p.n_blocks) * p.box_bytes;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        int st = 0;
        uint32_t ph = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            const int m0 = kb * p.pix;
            int img = 0, y0 = 0, x0 = 0;
            if (!p.linear) {
                img = m0 / p.HoWo;
                const int rem = m0 - img * p.HoWo;
                const int oy = rem / p.Wo;
                y0 = oy * p.stride - p.pad_h;
                x0 = (rem - oy * p.Wo) * p.stride - p.pad_w;
            }
            mbar_wait(&empty[st], ph ^ 1);
            if (elect_one()) {
                uint8_t* dst = ring + st * p.stage_bytes;
                mbar_arrive_expect_tx(&full[st], tx_bytes);
                tma_load_2d(&tmDy, &full[st], dst, co0, m0);
                if (a_boxes == 2) tma_load_2d(&tmDy, &full[st], dst + p.box_bytes, co0 + 64, m0);
                for (int g = 0; g < ntaps; ++g) {
                    const int tap = tap0 + g;
                    const int r = tap / p.kw, s = tap - r * p.kw;
                    for (int j = 0; j < p.n_blocks; ++j) {
                        uint8_t* b_dst = dst + (2 + g * p.n_blocks + j) * p.box_bytes;
                        if (p.linear) tma_load_2d(&tmX, &full[st], b_dst, ci0 + j * 64, m0);
                        else
                            tma_load_im2col_4d(&tmX, &full[st], b_dst, ci0 + j * 64, x0, y0, img, static_cast<uint16_t>(s),
                                               static_cast<uint16_t>(r));
                    }
                }
            }
            __syncwarp();
            if (++st == p.stages) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =====================================
        int st = 0;
        uint32_t ph = 0;
        // MN-major SW128 descriptors: SBO = 1024 B between 8-pixel groups, LBO = one [P px][64 ch] box between
        // 64-channel blocks; stepping 16 pixels along K = +2048 B on the start address
        const uint32_t dhi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
        const uint32_t lbo = ((p.box_bytes >> 4) & 0x3FFFu) << 16;
        const uint32_t ring16 = (smem_u32(ring) >> 4) & 0x3FFFu;
        const uint32_t box16 = p.box_bytes >> 4;
        const int ksteps = p.pix / 16;
        uint32_t accum = 0;
        for (int i = 0; i < nkb; ++i) {
            mbar_wait(&full[st], ph);
            tc_fence_after();
            const uint32_t a16 = ring16 + ((st * p.stage_bytes) >> 4);
            if (elect_one()) {
                for (int g = 0; g < ntaps; ++g) {
                    const uint32_t b16 = a16 + (2 + g * p.n_blocks) * box16;
                    const uint32_t d = tmem_base + g * bn;
                    for (int k = 0; k < ksteps; ++k)
                        umma_f16_ss_lohi(d, (a16 + k * 128) | lbo, (b16 + k * 128) | lbo, dhi, p.idesc, (accum | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty[st]);
                if (i == nkb - 1) umma_commit(acc_full);
            }
            __syncwarp();
            accum = 1;
            if (++st == p.stages) { st = 0; ph ^= 1; }
        }
    } else if (nkb > 0) {
        // ===================================== epilogue: TMEM -> fp32 reductions =====================================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int co = co0 + q * 32 + lane;
        const bool row_ok = co < p.Cout;
        const bool vec_ok = (p.Cin & 3) == 0;
        for (int g = 0; g < ntaps; ++g) {
            float* row = p.dw + (static_cast<size_t>(row_ok ? co : 0) * p.taps + tap0 + g) * p.Cin;
            for (int c = 0; c < p.n_blocks * 2; ++c) {
                const int cbase = ci0 + c * 32;
                if (cbase >= p.Cin) break;  // warp-uniform
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * bn + c * 32, v);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int ci = cbase + 4 * j;
                        if (vec_ok && ci + 3 < p.Cin) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + ci), "f"(__uint_as_float(v[4 * j])),
                                         "f"(__uint_as_float(v[4 * j + 1])), "f"(__uint_as_float(v[4 * j + 2])),
                                         "f"(__uint_as_float(v[4 * j + 3]))
                                         : "memory");
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (ci + e < p.Cin) atomicAdd(row + ci + e, __uint_as_float(v[4 * j + e]));
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

}  // namespace y5

using namespace y5;

extern "C" Y5_API int y5_conv_wgrad(const y5_wgrad_desc* d, void* stream) {
    if (!d || !d->in || !d->dout || !d->dweight) return set_error(Y5_E_INVALID, "wgrad: null pointer");
    if (d->dtype != Y5_F16 && d->dtype != Y5_BF16) return set_error(Y5_E_UNSUPPORTED, "wgrad: dtype must be fp16 or bf16");
    if (d->batch <= 0 || d->in_h <= 0 || d->in_w <= 0 || d->in_c <= 0 || d->out_c <= 0 || d->ksize <= 0 || d->stride <= 0 || d->pad < 0)
        return set_error(Y5_E_INVALID, "wgrad: bad shape");
    if ((d->in_pitch % 8) || (d->dout_pitch % 8) || (reinterpret_cast<uintptr_t>(d->in) & 15) || (reinterpret_cast<uintptr_t>(d->dout) & 15) ||
        (reinterpret_cast<uintptr_t>(d->dweight) & 15))
        return set_error(Y5_E_INVALID, "wgrad: views must be 16-byte aligned with pitches that are multiples of 8 elements");
    const bool strided_view = d->in_x_stride || d->in_y_stride || d->in_n_stride;
    if ((!strided_view && d->in_pitch < d->in_c) || d->dout_pitch < d->out_c)
        return set_error(Y5_E_INVALID, "wgrad: pitch smaller than channel count");
    if ((d->in_x_stride % 8) || (d->in_y_stride % 8) || (d->in_n_stride % 8) || d->kw < 0 || d->pad_w < 0)
        return set_error(Y5_E_INVALID, "wgrad: strides must be multiples of 8 elements");
    const int k = d->ksize;                       // filter rows
    const int kw = d->kw ? d->kw : d->ksize;      // filter columns
    const int pad_w = d->kw ? d->pad_w : d->pad;
    const int Ho = (d->in_h + 2 * d->pad - k) / d->stride + 1, Wo = (d->in_w + 2 * pad_w - kw) / d->stride + 1;
    if (Ho <= 0 || Wo <= 0) return set_error(Y5_E_INVALID, "wgrad: empty output");
    const long long M = static_cast<long long>(d->batch) * Ho * Wo;
    if (M > 0x7fffffffLL) return set_error(Y5_E_UNSUPPORTED, "wgrad: too many pixels");
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    WgradParams p{};
    p.M = static_cast<int>(M);
    p.Cout = d->out_c;
    p.Cin = d->in_c;
    p.kh = k;
    p.kw = kw;
    p.stride = d->stride;
    p.pad_h = d->pad;
    p.pad_w = pad_w;
    p.Wo = Wo;
    p.HoWo = Ho * Wo;
    p.linear = (k == 1 && kw == 1 && d->stride == 1 && d->pad == 0 && pad_w == 0 && !strided_view) ? 1 : 0;
    p.taps = k * kw;
    const int blocks_total = (d->in_c + 63) / 64;
    p.n_blocks = blocks_total < 4 ? blocks_total : 4;
    p.ci_tiles = (blocks_total + p.n_blocks - 1) / p.n_blocks;
    // balance the blocks over the ci tiles (e.g. 5 blocks -> 3 + 2, not 4 + 1)
    p.n_blocks = (blocks_total + p.ci_tiles - 1) / p.ci_tiles;
    const int co_tiles = (d->out_c + kWgCo - 1) / kWgCo;
    const int bn = p.n_blocks * 64;
    // taps per CTA: accumulators of a group share TMEM (512 columns); a 256-wide tile keeps one tap (its stage is
    // already 48 KB); groups are balanced (9 taps, limit 5 -> 5 + 4)
    static const int g_cap = [] { const char* e = getenv("Y5_WG_GROUP_MAX"); return e && atoi(e) > 0 ? atoi(e) : 5; }();
    static const unsigned stage_kb = [] { const char* e = getenv("Y5_WG_STAGE_KB"); return e && atoi(e) > 0 ? (unsigned)atoi(e) : 56u; }();
    int gmax = bn >= 256 ? 1 : 512 / bn;
    if (gmax > g_cap) gmax = g_cap;
    p.tap_groups = (p.taps + gmax - 1) / gmax;
    p.group = (p.taps + p.tap_groups - 1) / p.tap_groups;
    // pixels per stage: as many as keep a stage within ~56 KB (>= 3 stages in flight)
    p.pix = 64;
    for (int cand : {256, 128}) {
        if (static_cast<uint32_t>(2 + p.group * p.n_blocks) * cand * 128u <= stage_kb * 1024u) { p.pix = cand; break; }
    }
    p.box_bytes = p.pix * 128u;
    p.kblocks = (p.M + p.pix - 1) / p.pix;
    const long long items = static_cast<long long>(co_tiles) * p.tap_groups * p.ci_tiles;
    long long want = (sm_count() + items - 1) / items;  // pixel ranges per tile: one CTA per SM, one wave
    if (want < 1) want = 1;
    if (want > p.kblocks) want = p.kblocks;
    p.kb_per_split = static_cast<int>((p.kblocks + want - 1) / want);
    p.splits = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty range
    p.stage_bytes = (2 + p.group * p.n_blocks) * p.box_bytes;
    p.stages = static_cast<int>((200u * 1024u) / p.stage_bytes);
    if (p.stages > kWgStagesMax) p.stages = kWgStagesMax;
    if (p.stages < 2) return set_error(Y5_E_UNSUPPORTED, "wgrad: stage does not fit shared memory");
    p.idesc = umma_idesc_f16(d->dtype == Y5_BF16, bn) | (1u << 15) | (1u << 16);  // A and B MN-major
    const int cols = p.group * bn;
    p.tmem_cols = cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512));
    p.dw = d->dweight;

    CUtensorMap tmDy, tmX;
    {
        cuuint64_t dims[2] = {(cuuint64_t)d->out_c, (cuuint64_t)M};
        cuuint64_t str[1] = {(cuuint64_t)d->dout_pitch * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)p.pix};
        int e = encode_tiled(&tmDy, d->dtype, d->dout, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, "wgrad dY");
        if (e) return e;
    }
    if (p.linear) {
        cuuint64_t dims[2] = {(cuuint64_t)d->in_c, (cuuint64_t)M};
        cuuint64_t str[1] = {(cuuint64_t)d->in_pitch * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)p.pix};
        int e = encode_tiled(&tmX, d->dtype, d->in, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, "wgrad X");
        if (e) return e;
    } else {
        const long long xs = d->in_x_stride ? d->in_x_stride : d->in_pitch;
        const long long ys = d->in_y_stride ? d->in_y_stride : xs * d->in_w;
        const long long ns = d->in_n_stride ? d->in_n_stride : ys * d->in_h;
        int e = encode_im2col(&tmX, d->dtype, d->in, d->in_c, d->in_w, d->in_h, d->batch, xs, ys, ns, k, kw, d->stride, d->pad, pad_w, 64,
                              p.pix, CU_TENSOR_MAP_SWIZZLE_128B);
        if (e) return e;
    }
    const size_t dw_bytes = static_cast<size_t>(d->out_c) * p.taps * d->in_c * sizeof(float);
    if (!d->accumulate) {
        cudaError_t me = cudaMemsetAsync(d->dweight, 0, dw_bytes, st);
        if (me != cudaSuccess) return set_error(int(me), "wgrad: memset failed: %s", cudaGetErrorString(me));
    }
    const uint32_t smem = p.stages * p.stage_bytes + (2 * kWgStagesMax + 1) * 8 + 16 + 1024;
    const cudaError_t attr_err = ensure_dyn_smem(reinterpret_cast<const void*>(conv_wgrad_kernel), 227 * 1024);
    if (attr_err != cudaSuccess) return set_error(int(attr_err), "wgrad: cudaFuncSetAttribute failed");
    const long long grid = items * p.splits;
    count_launch();
    launch_pdl(conv_wgrad_kernel, dim3(static_cast<unsigned>(grid)), dim3(kWgThreads), smem, st, tmDy, tmX, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(int(e), "wgrad launch failed: %s", cudaGetErrorString(e));
    return 0;
}
