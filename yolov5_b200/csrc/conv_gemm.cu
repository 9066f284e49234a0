// Implicit-GEMM convolution for sm_100a:  D[M = B*Ho*Wo, N = Cout] = im2col(X)[M, K = kh*kw*Cin] * W[N, K]^T
//
//   * A (activations, NHWC view) is fetched by TMA -- plain 2-D tiles for 1x1/s1 convs, hardware im2col mode
//     (cp.async.bulk.tensor.4d...im2col) for everything else, so a 128-row tile is 128 consecutive output pixels
//     in (n, y, x) order regardless of row / image boundaries and padding is zero-filled by the copy engine;
//   * B (weights, [Cout][kh][kw][Cin_pad], BN scale folded in) by 2-D TMA tiles; both land in 32/64/128-byte
//     swizzled K-major shared-memory tiles that tcgen05.mma consumes directly;
//   * one elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into a TMEM accumulator (2 accumulators, so
//     the epilogue of tile i overlaps the MMAs of tile i+1);
//   * epilogue warps read TMEM with tcgen05.ld, add the folded-BN bias, apply SiLU, add the optional residual and
//     store fp16/bf16 NHWC into a (possibly strided) channel slice -- or, for the Detect head, write the raw
//     (B,na,ny,nx,no) logits and the decoded (B,rows,no) predictions.
//   Persistent grid (one CTA per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner,
//   warps 2..5 epilogue.
//
// Replaces reference models/common.py:86-92 (Conv), :181 (Bottleneck add), :246/:340/:453 (cat, via strided
// output views) and models/yolo.py:95-113 (Detect level).
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/y5b200.h"
#include "common.cuh"
#include "host_util.h"

namespace y5 {

constexpr int kBlockM = 128;
constexpr int kThreads = 192;      // warp0 producer, warp1 mma, warps2-5 epilogue
constexpr int kEpiThreads = 128;
constexpr int kMaxStages = 8;
constexpr int kDetectStagePitch = 258;  // halves per staged row (256 + 2: odd word pitch -> conflict-free)

struct ConvParams {
    int M, N;                  // GEMM rows (B*Ho*Wo), output channels
    int num_m_tiles, num_n_tiles;
    int kh, kw, c_chunks;      // K loop = kh*kw*c_chunks blocks of block_k
    int block_k;               // 16 | 32 | 64 elements (row bytes 32/64/128)
    int a_im2col;              // 0: 2-D tiled A, 1: im2col A
    int Wo, HoWo, stride, pad;
    int num_stages;
    uint32_t a_stage_bytes, b_stage_bytes;
    uint32_t idesc;
    int is_bf16, act;
    const float* bias;
    // EPI 0
    void* out;
    int out_pitch;
    const void* res;
    int res_pitch;
    // EPI 1 (detect)
    void* raw;
    void* z;
    int na, no, nc, nx, z_rows, z_row0;
    float det_stride;
    float anchor_wh[8];
};

struct SmemLayout {
    uint32_t off_a, off_b, off_bias, off_stage, off_lut, off_rowinfo, off_bars, off_tmem, total;
};

__host__ __device__ inline SmemLayout smem_layout(int block_n, int epi, int stages, uint32_t a_bytes, uint32_t b_bytes) {
    SmemLayout L;
    uint32_t o = 0;
    L.off_a = o;
    o += stages * a_bytes;
    L.off_b = o;
    o += stages * b_bytes;
    L.off_bias = o;
    o += block_n * 4;
    L.off_stage = o;
    if (epi == 1) o += kBlockM * kDetectStagePitch * 2;
    L.off_lut = o;
    if (epi == 1) o += 512 * 2 * 2;  // col -> anchor, col -> o (uint16 each), up to 512 columns
    L.off_rowinfo = o;
    if (epi == 1) o += kBlockM * 16;  // per row: int64 raw offset (a=0), int64 z offset (a=0); gx,gy recomputed
    o = (o + 7) & ~7u;
    L.off_bars = o;
    o += (2 * kMaxStages + 4) * 8;
    L.off_tmem = o;
    o += 16;
    L.total = o;
    return L;
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic smem is only guaranteed 16B aligned by the ABI; round up ourselves (host adds 1024 B of slack)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemLayout L = smem_layout(BLOCK_N, EPI, p.num_stages, p.a_stage_bytes, p.b_stage_bytes);
    uint8_t* sA = smem + L.off_a;
    uint8_t* sB = smem + L.off_b;
    float* sBias = reinterpret_cast<float*>(smem + L.off_bias);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.off_bars);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full = empty_bar + kMaxStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L.off_tmem);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.num_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, kTmemCols);
    if (EPI == 1 && warp >= 2) {  // column -> (anchor, o) lookup for the head epilogue
        uint16_t* lut = reinterpret_cast<uint16_t*>(smem + L.off_lut);
        for (int c = threadIdx.x - 64; c < 512; c += kEpiThreads) {
            lut[c] = static_cast<uint16_t>(c / p.no);
            lut[512 + c] = static_cast<uint16_t>(c % p.no);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int num_kb = p.kh * p.kw * p.c_chunks;
    const uint32_t row_bytes = p.block_k * 2;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = p.a_stage_bytes + p.b_stage_bytes;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / p.num_n_tiles) * kBlockM;
                const int n0 = (tile % p.num_n_tiles) * BLOCK_N;
                int img = 0, base_h = 0, base_w = 0;
                if (p.a_im2col) {
                    img = m0 / p.HoWo;
                    const int rem = m0 - img * p.HoWo;
                    const int oy = rem / p.Wo;
                    const int ox = rem - oy * p.Wo;
                    base_h = oy * p.stride - p.pad;
                    base_w = ox * p.stride - p.pad;
                }
                int kb = 0;
                for (int r = 0; r < p.kh; ++r) {
                    for (int s = 0; s < p.kw; ++s) {
                        for (int cc = 0; cc < p.c_chunks; ++cc, ++kb) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                            uint8_t* a_dst = sA + stage * p.a_stage_bytes;
                            uint8_t* b_dst = sB + stage * p.b_stage_bytes;
                            if (p.a_im2col)
                                tma_load_im2col_4d(&tmA, &full_bar[stage], a_dst, cc * p.block_k, base_w, base_h, img,
                                                   static_cast<uint16_t>(s), static_cast<uint16_t>(r));
                            else
                                tma_load_2d(&tmA, &full_bar[stage], a_dst, cc * p.block_k, m0);
                            tma_load_2d(&tmB, &full_bar[stage], b_dst, kb * p.block_k, n0);
                            if (++stage == p.num_stages) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =====================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const int k_steps = p.block_k / 16;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * p.a_stage_bytes);
                    const uint32_t b_addr = smem_u32(sB + stage * p.b_stage_bytes);
                    for (int k = 0; k < k_steps; ++k) {
                        const uint64_t ad = umma_smem_desc(a_addr + k * 32, row_bytes);
                        const uint64_t bd = umma_smem_desc(b_addr + k * 32, row_bytes);
                        umma_f16_ss(d_tmem, ad, bd, p.idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
                    if (++stage == p.num_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===================================== epilogue (warps 2..5) =====================================
        const int et = threadIdx.x - 64;         // 0..127
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;     // accumulator row (= TMEM lane) owned by this thread
        const bool bf16 = p.is_bf16 != 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / p.num_n_tiles) * kBlockM;
            const int n0 = (tile % p.num_n_tiles) * BLOCK_N;
            named_bar_sync(1, kEpiThreads);  // previous tile's sBias / staging reads are finished
            for (int i = et; i < BLOCK_N; i += kEpiThreads) sBias[i] = (n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.0f;
            if (EPI == 1) {
                long long* rowinfo = reinterpret_cast<long long*>(smem + L.off_rowinfo);
                const int m = m0 + et;
                long long ro = -1, zo = -1;
                if (m < p.M) {
                    const int b = m / p.HoWo;
                    const int pix = m - b * p.HoWo;
                    ro = (static_cast<long long>(b) * p.na * p.HoWo + pix) * p.no;
                    zo = (static_cast<long long>(b) * p.z_rows + p.z_row0 + pix) * p.no;
                }
                rowinfo[2 * et] = ro;
                rowinfo[2 * et + 1] = zo;
            }
            named_bar_sync(1, kEpiThreads);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
            const int m = m0 + row;
            const bool valid = m < p.M;

            if (EPI == 0) {
                uint8_t* out_row = reinterpret_cast<uint8_t*>(p.out) + static_cast<size_t>(m) * p.out_pitch * 2;
                const uint8_t* res_row =
                    p.res ? reinterpret_cast<const uint8_t*>(p.res) + static_cast<size_t>(m) * p.res_pitch * 2 : nullptr;
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_row + c * 32, v);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int col = c * 32 + g * 8;
                            if (n0 + col < p.N) {
                                float f[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    float x = __uint_as_float(v[g * 8 + j]) + sBias[col + j];
                                    f[j] = p.act ? silu_f(x) : x;
                                }
                                if (res_row) {
                                    const uint4 rv = *reinterpret_cast<const uint4*>(res_row + (n0 + col) * 2);
                                    const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const float2 t = unpack2(rr[j], bf16);
                                        f[2 * j] += t.x;
                                        f[2 * j + 1] += t.y;
                                    }
                                }
                                uint4 o;
                                o.x = pack2(f[0], f[1], bf16);
                                o.y = pack2(f[2], f[3], bf16);
                                o.z = pack2(f[4], f[5], bf16);
                                o.w = pack2(f[6], f[7], bf16);
                                *reinterpret_cast<uint4*>(out_row + (n0 + col) * 2) = o;
                            }
                        }
                    }
                }
            } else {
                // ---- Detect head: pass 0 writes raw logits, pass 1 the decoded predictions (models/yolo.py:95-113) ----
                uint16_t* stage_buf = reinterpret_cast<uint16_t*>(smem + L.off_stage);
                const uint16_t* lut_a = reinterpret_cast<const uint16_t*>(smem + L.off_lut);
                const uint16_t* lut_o = lut_a + 512;
                const long long* rowinfo = reinterpret_cast<const long long*>(smem + L.off_rowinfo);
                const int ncols = min(BLOCK_N, p.N - n0);
                int gx = 0, gy = 0;
                if (valid) {
                    const int pix = m % p.HoWo;
                    gy = pix / p.nx;
                    gx = pix - gy * p.nx;
                }
                for (int pass = 0; pass < 2; ++pass) {
                    uint16_t* srow = stage_buf + row * kDetectStagePitch;
#pragma unroll 1
                    for (int c = 0; c < BLOCK_N / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_row + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = c * 32 + j;
                            if (col < ncols) {
                                float x = __uint_as_float(v[j]) + sBias[col];
                                if (pass == 1) {
                                    const int a = lut_a[n0 + col], o = lut_o[n0 + col];
                                    if (o < 5 + p.nc) {
                                        const float s = sigmoid_f(x);
                                        if (o == 0) x = (s * 2.0f + (static_cast<float>(gx) - 0.5f)) * p.det_stride;
                                        else if (o == 1) x = (s * 2.0f + (static_cast<float>(gy) - 0.5f)) * p.det_stride;
                                        else if (o < 4) { const float t = s * 2.0f; x = t * t * p.anchor_wh[a * 2 + (o - 2)]; }
                                        else x = s;
                                    }
                                }
                                srow[col] = pack1(x, bf16);
                            }
                        }
                    }
                    named_bar_sync(1, kEpiThreads);
                    uint16_t* gout = reinterpret_cast<uint16_t*>(pass == 0 ? p.raw : p.z);
                    const long long a_stride = static_cast<long long>(p.HoWo) * p.no;
                    const int total = kBlockM * ncols;
                    for (int e = et; e < total; e += kEpiThreads) {
                        const int r = e / ncols;
                        const int col = e - r * ncols;
                        const long long base = rowinfo[2 * r + pass];
                        if (base >= 0) {
                            const int a = lut_a[n0 + col], o = lut_o[n0 + col];
                            gout[base + a * a_stride + o] = stage_buf[r * kDetectStagePitch + col];
                        }
                    }
                    named_bar_sync(1, kEpiThreads);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Independent CUDA-core direct convolution (cross-check on device; also documents the packed weight layout).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void conv_direct_kernel(const uint16_t* __restrict__ in, int in_pitch, int B, int H, int W, int Cin,
                                   const uint16_t* __restrict__ w, int cin_pad, const float* __restrict__ bias,
                                   uint16_t* out, int out_pitch, int Cout, const uint16_t* res, int res_pitch, int ks,
                                   int stride, int pad, int Ho, int Wo, int act, int bf16) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(B) * Ho * Wo * Cout;
    if (idx >= total) return;
    const int n = static_cast<int>(idx % Cout);
    const long long m = idx / Cout;
    const int ox = static_cast<int>(m % Wo);
    const int oy = static_cast<int>((m / Wo) % Ho);
    const int b = static_cast<int>(m / (static_cast<long long>(Wo) * Ho));
    float acc = 0.0f;
    for (int r = 0; r < ks; ++r) {
        const int iy = oy * stride - pad + r;
        if (iy < 0 || iy >= H) continue;
        for (int s = 0; s < ks; ++s) {
            const int ix = ox * stride - pad + s;
            if (ix < 0 || ix >= W) continue;
            const uint16_t* ip = in + (static_cast<long long>(b) * H * W + static_cast<long long>(iy) * W + ix) * in_pitch;
            const uint16_t* wp = w + (static_cast<long long>(n) * ks * ks + r * ks + s) * cin_pad;
            for (int c = 0; c < Cin; ++c) acc += unpack1(ip[c], bf16) * unpack1(wp[c], bf16);
        }
    }
    float x = acc + bias[n];
    if (act) x = x / (1.0f + expf(-x));
    if (res) x += unpack1(res[m * res_pitch + n], bf16);
    out[m * out_pitch + n] = pack1(x, bf16);
}

}  // namespace y5

// =====================================================================================================================
// host side
// =====================================================================================================================
using namespace y5;

namespace {

int pick_block_k(int in_c) {
    int best = 64;
    long best_cost = -1;
    for (int bk : {64, 32, 16}) {
        const int chunks = (in_c + bk - 1) / bk;
        const long cost = static_cast<long>(chunks) * (bk + 24);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = bk;
        }
    }
    return best;
}

int pick_block_n(int out_c, int64_t m_rows) {
    if (out_c <= 32) return 32;
    if (out_c <= 64) return 64;
    if (out_c <= 128) return 128;
    const int64_t m_tiles = (m_rows + kBlockM - 1) / kBlockM;
    // prefer 256-wide tiles (half the B-operand smem traffic per flop) when that still leaves >= 2 waves of tiles
    const int64_t tiles256 = m_tiles * ((out_c + 255) / 256);
    if (tiles256 >= 2 * 148 && (out_c % 256 == 0 || out_c > 384)) return 256;
    return 128;
}

struct TmapSpec {
    CUtensorMap map;
};

CUtensorMapSwizzle swizzle_for(int block_k) {
    return block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (block_k == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

int encode_2d(CUtensorMap* map, int dtype, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
              uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
    auto fn = driver_fn_encode_tiled();
    if (!fn) return set_error(Y5_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {outer_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dtype == Y5_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(Y5_E_DRIVER, "cuTensorMapEncodeTiled failed (%d): inner %llu outer %llu stride %llu box %ux%u",
                                            int(r), (unsigned long long)inner, (unsigned long long)outer,
                                            (unsigned long long)outer_stride_bytes, box_inner, box_outer);
    return 0;
}

int encode_im2col(CUtensorMap* map, int dtype, const void* base, int C, int W, int H, int N, int pitch, int ks, int stride,
                  int pad, uint32_t channels_per_pixel, uint32_t pixels_per_column, CUtensorMapSwizzle sw) {
    auto fn = driver_fn_encode_im2col();
    if (!fn) return set_error(Y5_E_DRIVER, "cuTensorMapEncodeIm2col entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2};
    // Bounding box of filter-window base pixels: lower corner = -pad, upper corner = pad - (k-1) (dilation 1),
    // relative to the tensor's first / last pixel; the window taps {s, r} are passed per copy as im2col offsets.
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - (ks - 1), pad - (ks - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = fn(map, dtype == Y5_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                    const_cast<void*>(base), dims, strides, lower, upper, channels_per_pixel, pixels_per_column, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(Y5_E_DRIVER, "cuTensorMapEncodeIm2col failed (%d): C %d W %d H %d N %d pitch %d k %d s %d p %d", int(r), C, W,
                         H, N, pitch, ks, stride, pad);
    // Driver-side quirk also worked around by CUTLASS (cute/atom/copy_traits_sm90_im2col.hpp): for tensors smaller
    // than 128 KiB, drivers <= 13.1 set a descriptor bit that makes the im2col walk fault; clear it.
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const unsigned long long span = (unsigned long long)N * H * W * pitch * 2;
    if (drv <= 13010 && span < 131072ull) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
    return 0;
}

template <int BN, int EPI>
cudaError_t launch_conv(const CUtensorMap& a, const CUtensorMap& b, const ConvParams& p, int grid, uint32_t smem, cudaStream_t st) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(conv_gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess) return attr_err;
    conv_gemm_kernel<BN, EPI><<<grid, kThreads, smem, st>>>(a, b, p);
    count_launch();
    return cudaGetLastError();
}

struct PlanCommon {
    CUtensorMap tmA, tmB;
    ConvParams p;
    int block_n, epi, grid;
    uint32_t smem_bytes;
};

int finish_plan(PlanCommon& pc, int in_c, int block_k, int block_n, int epi) {
    ConvParams& p = pc.p;
    p.block_k = block_k;
    p.c_chunks = (in_c + block_k - 1) / block_k;
    p.a_stage_bytes = kBlockM * block_k * 2;
    p.b_stage_bytes = block_n * block_k * 2;
    p.num_m_tiles = (p.M + kBlockM - 1) / kBlockM;
    p.num_n_tiles = (p.N + block_n - 1) / block_n;
    p.idesc = umma_idesc_f16(p.is_bf16 != 0, block_n);
    const uint32_t budget = 226 * 1024 - 1024;
    int stages = kMaxStages;
    for (; stages >= 2; --stages) {
        if (smem_layout(block_n, epi, stages, p.a_stage_bytes, p.b_stage_bytes).total <= budget) break;
    }
    if (stages < 2) return set_error(Y5_E_UNSUPPORTED, "conv tile does not fit shared memory");
    const int num_kb = p.kh * p.kw * p.c_chunks;
    if (stages > num_kb + 1) stages = num_kb + 1 < 2 ? 2 : num_kb + 1;
    p.num_stages = stages;
    pc.smem_bytes = smem_layout(block_n, epi, stages, p.a_stage_bytes, p.b_stage_bytes).total + 1024;
    pc.block_n = block_n;
    pc.epi = epi;
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    const int sms = sm_count();
    pc.grid = tiles < sms ? tiles : sms;
    return 0;
}

int run_plan(const PlanCommon& pc, cudaStream_t st) {
    cudaError_t e = cudaErrorInvalidValue;
    if (pc.epi == 0) {
        switch (pc.block_n) {
            case 32: e = launch_conv<32, 0>(pc.tmA, pc.tmB, pc.p, pc.grid, pc.smem_bytes, st); break;
            case 64: e = launch_conv<64, 0>(pc.tmA, pc.tmB, pc.p, pc.grid, pc.smem_bytes, st); break;
            case 128: e = launch_conv<128, 0>(pc.tmA, pc.tmB, pc.p, pc.grid, pc.smem_bytes, st); break;
            case 256: e = launch_conv<256, 0>(pc.tmA, pc.tmB, pc.p, pc.grid, pc.smem_bytes, st); break;
        }
    } else {
        e = launch_conv<256, 1>(pc.tmA, pc.tmB, pc.p, pc.grid, pc.smem_bytes, st);
    }
    if (e != cudaSuccess) return set_error(int(e), "conv_gemm launch failed: %s", cudaGetErrorString(e));
    return 0;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

struct y5_conv_plan {
    PlanCommon pc;
};
struct y5_detect_plan {
    PlanCommon pc;
};

extern "C" Y5_API int y5_conv_pick(int32_t in_c, int32_t out_c, int64_t m_rows, int32_t* block_k, int32_t* block_n) {
    if (in_c <= 0 || out_c <= 0) return set_error(Y5_E_INVALID, "y5_conv_pick: non-positive channel count");
    if (block_k) *block_k = pick_block_k(in_c);
    if (block_n) *block_n = pick_block_n(out_c, m_rows);
    return 0;
}

static int validate_conv(const y5_conv_desc* d) {
    if (!d || !d->in || !d->weight || !d->bias || !d->out) return set_error(Y5_E_INVALID, "conv: null pointer");
    if (d->dtype != Y5_F16 && d->dtype != Y5_BF16) return set_error(Y5_E_UNSUPPORTED, "conv: dtype must be fp16/bf16");
    if (d->batch <= 0 || d->in_h <= 0 || d->in_w <= 0 || d->in_c <= 0 || d->out_c <= 0)
        return set_error(Y5_E_INVALID, "conv: non-positive dimension");
    if (d->in_pitch < d->in_c || d->in_pitch % 8 || d->out_pitch < d->out_c || d->out_pitch % 8 || d->out_c % 8)
        return set_error(Y5_E_INVALID, "conv: pitches/out_c must be multiples of 8 elements and cover the view");
    if (!aligned16(d->in) || !aligned16(d->out) || !aligned16(d->weight) || (d->residual && !aligned16(d->residual)))
        return set_error(Y5_E_INVALID, "conv: pointers must be 16-byte aligned");
    if (d->residual && (d->res_pitch < d->out_c || d->res_pitch % 8)) return set_error(Y5_E_INVALID, "conv: bad residual pitch");
    if (d->ksize < 1 || d->ksize > 7 || d->stride < 1 || d->stride > 8 || d->pad < 0 || d->pad > d->ksize)
        return set_error(Y5_E_UNSUPPORTED, "conv: kernel %d stride %d pad %d unsupported", d->ksize, d->stride, d->pad);
    return 0;
}

extern "C" Y5_API int y5_conv_plan_create(const y5_conv_desc* d, y5_conv_plan** out) {
    if (!out) return set_error(Y5_E_INVALID, "conv: null plan out");
    *out = nullptr;
    if (int e = validate_conv(d)) return e;
    const int Ho = (d->in_h + 2 * d->pad - d->ksize) / d->stride + 1;
    const int Wo = (d->in_w + 2 * d->pad - d->ksize) / d->stride + 1;
    if (Ho <= 0 || Wo <= 0) return set_error(Y5_E_INVALID, "conv: empty output");
    const int64_t M64 = static_cast<int64_t>(d->batch) * Ho * Wo;
    if (M64 > 0x7fffffff - 256) return set_error(Y5_E_UNSUPPORTED, "conv: more than 2^31 output pixels");
    int bk = d->block_k ? d->block_k : pick_block_k(d->in_c);
    int bn = d->block_n ? d->block_n : pick_block_n(d->out_c, M64);
    if (bk != 16 && bk != 32 && bk != 64) return set_error(Y5_E_INVALID, "conv: block_k must be 16/32/64");
    if (bn != 32 && bn != 64 && bn != 128 && bn != 256) return set_error(Y5_E_INVALID, "conv: block_n must be 32/64/128/256");
    auto* plan = new y5_conv_plan();
    PlanCommon& pc = plan->pc;
    std::memset(&pc.p, 0, sizeof(pc.p));
    ConvParams& p = pc.p;
    p.M = static_cast<int>(M64);
    p.N = d->out_c;
    p.kh = p.kw = d->ksize;
    p.Wo = Wo;
    p.HoWo = Ho * Wo;
    p.stride = d->stride;
    p.pad = d->pad;
    p.is_bf16 = d->dtype == Y5_BF16;
    p.act = d->act;
    p.bias = d->bias;
    p.out = d->out;
    p.out_pitch = d->out_pitch;
    p.res = d->residual;
    p.res_pitch = d->res_pitch;
    p.a_im2col = !(d->ksize == 1 && d->stride == 1 && d->pad == 0);
    if (int e = finish_plan(pc, d->in_c, bk, bn, 0)) { delete plan; return e; }
    const CUtensorMapSwizzle sw = swizzle_for(bk);
    int e;
    if (p.a_im2col)
        e = encode_im2col(&pc.tmA, d->dtype, d->in, d->in_c, d->in_w, d->in_h, d->batch, d->in_pitch, d->ksize, d->stride, d->pad,
                          bk, kBlockM, sw);
    else
        e = encode_2d(&pc.tmA, d->dtype, d->in, d->in_c, static_cast<uint64_t>(p.M), static_cast<uint64_t>(d->in_pitch) * 2, bk,
                      kBlockM, sw);
    if (e) { delete plan; return e; }
    const uint64_t ktot = static_cast<uint64_t>(p.kh) * p.kw * p.c_chunks * bk;
    e = encode_2d(&pc.tmB, d->dtype, d->weight, ktot, d->out_c, ktot * 2, bk, bn, sw);
    if (e) { delete plan; return e; }
    *out = plan;
    return 0;
}

extern "C" Y5_API int y5_conv_plan_run(const y5_conv_plan* plan, void* stream) {
    if (!plan) return set_error(Y5_E_INVALID, "conv: null plan");
    return run_plan(plan->pc, static_cast<cudaStream_t>(stream));
}
extern "C" Y5_API void y5_conv_plan_destroy(y5_conv_plan* plan) { delete plan; }

extern "C" Y5_API int y5_conv_bn_silu_fwd(const y5_conv_desc* d, void* stream) {
    y5_conv_plan* plan = nullptr;
    if (int e = y5_conv_plan_create(d, &plan)) return e;
    const int e = y5_conv_plan_run(plan, stream);
    y5_conv_plan_destroy(plan);
    return e;
}

extern "C" Y5_API int y5_conv_direct_fwd(const y5_conv_desc* d, void* stream) {
    if (int e = validate_conv(d)) return e;
    const int Ho = (d->in_h + 2 * d->pad - d->ksize) / d->stride + 1;
    const int Wo = (d->in_w + 2 * d->pad - d->ksize) / d->stride + 1;
    const int bk = d->block_k ? d->block_k : pick_block_k(d->in_c);
    const int cin_pad = (d->in_c + bk - 1) / bk * bk;
    const long long total = static_cast<long long>(d->batch) * Ho * Wo * d->out_c;
    const int threads = 256;
    const long long blocks = (total + threads - 1) / threads;
    conv_direct_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(d->in), d->in_pitch, d->batch, d->in_h, d->in_w, d->in_c,
        static_cast<const uint16_t*>(d->weight), cin_pad, d->bias, static_cast<uint16_t*>(d->out), d->out_pitch, d->out_c,
        static_cast<const uint16_t*>(d->residual), d->res_pitch, d->ksize, d->stride, d->pad, Ho, Wo, d->act, d->dtype == Y5_BF16);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(int(e), "conv_direct launch failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" Y5_API int y5_detect_plan_create(const y5_detect_desc* d, y5_detect_plan** out) {
    if (!out) return set_error(Y5_E_INVALID, "detect: null plan out");
    *out = nullptr;
    if (!d || !d->in || !d->weight || !d->bias || !d->raw || !d->z) return set_error(Y5_E_INVALID, "detect: null pointer");
    if (d->dtype != Y5_F16 && d->dtype != Y5_BF16) return set_error(Y5_E_UNSUPPORTED, "detect: dtype must be fp16/bf16");
    if (d->na < 1 || d->na > 4 || d->no < 6 || d->na * d->no > 512 || d->nc < 1 || 5 + d->nc > d->no)
        return set_error(Y5_E_UNSUPPORTED, "detect: na %d no %d nc %d unsupported", d->na, d->no, d->nc);
    if (d->in_pitch < d->in_c || d->in_pitch % 8 || !aligned16(d->in) || !aligned16(d->weight))
        return set_error(Y5_E_INVALID, "detect: bad input view");
    const int64_t M64 = static_cast<int64_t>(d->batch) * d->ny * d->nx;
    if (M64 <= 0 || M64 > 0x7fffffff - 256) return set_error(Y5_E_INVALID, "detect: bad pixel count");
    const int bk = d->block_k ? d->block_k : pick_block_k(d->in_c);
    auto* plan = new y5_detect_plan();
    PlanCommon& pc = plan->pc;
    std::memset(&pc.p, 0, sizeof(pc.p));
    ConvParams& p = pc.p;
    p.M = static_cast<int>(M64);
    p.N = d->na * d->no;
    p.kh = p.kw = 1;
    p.Wo = d->nx;
    p.HoWo = d->ny * d->nx;
    p.stride = 1;
    p.pad = 0;
    p.is_bf16 = d->dtype == Y5_BF16;
    p.act = 0;
    p.bias = d->bias;
    p.raw = d->raw;
    p.z = d->z;
    p.na = d->na;
    p.no = d->no;
    p.nc = d->nc;
    p.nx = d->nx;
    p.z_rows = d->z_rows;
    p.z_row0 = d->z_row0;
    p.det_stride = d->stride;
    for (int i = 0; i < 8; ++i) p.anchor_wh[i] = d->anchor_wh[i];
    p.a_im2col = 0;
    if (int e = finish_plan(pc, d->in_c, bk, 256, 1)) { delete plan; return e; }
    const CUtensorMapSwizzle sw = swizzle_for(bk);
    int e = encode_2d(&pc.tmA, d->dtype, d->in, d->in_c, static_cast<uint64_t>(p.M), static_cast<uint64_t>(d->in_pitch) * 2, bk,
                      kBlockM, sw);
    if (e) { delete plan; return e; }
    const uint64_t ktot = static_cast<uint64_t>(p.c_chunks) * bk;
    e = encode_2d(&pc.tmB, d->dtype, d->weight, ktot, p.N, ktot * 2, bk, 256, sw);
    if (e) { delete plan; return e; }
    *out = plan;
    return 0;
}
extern "C" Y5_API int y5_detect_plan_run(const y5_detect_plan* plan, void* stream) {
    if (!plan) return set_error(Y5_E_INVALID, "detect: null plan");
    return run_plan(plan->pc, static_cast<cudaStream_t>(stream));
}
extern "C" Y5_API int y5_detect_plan_run_to(const y5_detect_plan* plan, void* raw, void* z, void* stream) {
    if (!plan || !raw || !z) return set_error(Y5_E_INVALID, "detect: null plan/output");
    PlanCommon pc = plan->pc;
    pc.p.raw = raw;
    pc.p.z = z;
    return run_plan(pc, static_cast<cudaStream_t>(stream));
}
extern "C" Y5_API void y5_detect_plan_destroy(y5_detect_plan* plan) { delete plan; }
