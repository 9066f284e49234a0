/* y5b200.h -- C ABI of liby5b200.so: the B200 (sm_100a) engine behind the YOLOv5 forward / NMS / loss hot path.
 *
 * The reference (ultralytics/yolov5) is pure Python and has NO FFI / plugin interface for this path: the work
 * sits behind ordinary Python callables.  Each entry point below therefore names the
 * reference callable whose body it replaces (file:line under /root/reference); INTEGRATION.md shows the ctypes
 * stubs a maintainer adds to those files.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; device pointers are raw CUDA device addresses; `stream` is a cudaStream_t;
 *   - never allocate device memory, never synchronise the stream, never touch the default stream;
 *   - activations are NHWC ("pixel-major"): element (n,y,x,c) of a view lives at base[((n*H+y)*W+x)*pitch + c],
 *     where `pitch` (elements per pixel of the underlying buffer) >= the view's channel count -- a view may be a
 *     channel slice of a wider buffer (this is how torch.cat is eliminated); pitches and channel offsets are
 *     multiples of 8 elements (16 bytes);
 *   - dtype: Y5_F16 or Y5_BF16 for activations and packed weights; accumulation and bias are fp32;
 *   - return 0 on success, a negative Y5_E* code for a rejected argument, a positive cudaError_t for a CUDA
 *     failure; y5_last_error() returns a thread-local message for the last non-zero return.
 */
#ifndef Y5B200_H
#define Y5B200_H
#include <stdint.h>

#if defined(__GNUC__)
#define Y5_API __attribute__((visibility("default")))
#else
#define Y5_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define Y5_F16 0
#define Y5_BF16 1
#define Y5_F32 2
#define Y5_U8 3

#define Y5_E_INVALID (-1)     /* bad argument (null pointer, misaligned pitch, ...) */
#define Y5_E_UNSUPPORTED (-2) /* shape outside what the kernels implement */
#define Y5_E_DRIVER (-3)      /* tensor-map encode / driver entry point unavailable */

#define Y5_ACT_NONE 0
#define Y5_ACT_SILU 1

int y5_version(void);
const char* y5_last_error(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
int64_t y5_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused Conv2d(bias=False) + folded BatchNorm + SiLU (+ residual add), implicit GEMM on tcgen05 tensor cores.
 * Replaces: models/common.py:86-92 Conv.forward / forward_fuse (conv -> bn -> act), utils/torch_utils.py:224-254
 * (BN fold, done once by the caller when packing), models/common.py:181 Bottleneck's `x + ...` (residual),
 * and, through out pitch/offset, the torch.cat of models/common.py:246,340,453.
 *   weights : packed [out_c][kh][kw][cin_pad] (K-major), cin_pad = chunks*block_k as reported by y5_conv_pick,
 *             zero padded, dtype = activation dtype, BN scale already folded in
 *   bias    : fp32 [out_c] (folded BN shift)
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct y5_conv_desc {
    const void* in;       /* view base (already offset to its first channel) */
    int32_t in_pitch;     /* elements per pixel of the buffer holding the view */
    int32_t batch, in_h, in_w, in_c;
    const void* weight;
    const float* bias;
    void* out;            /* view base (already offset to its first channel) */
    int32_t out_pitch;
    int32_t out_c;
    const void* residual; /* optional: same shape as out; may alias out (in-place add) */
    int32_t res_pitch;
    int32_t ksize, stride, pad;
    int32_t act;          /* Y5_ACT_* */
    int32_t dtype;        /* Y5_F16 | Y5_BF16 */
    int32_t block_k;      /* 16|32|64, or 0 = y5_conv_pick's choice; must match the weight packing */
    int32_t block_n;      /* 32|64|128|256, or 0 = auto */
    /* optional generalisations (all 0 = the plain case above) */
    int32_t kw;           /* non-square filter: width (height stays `ksize`); 0 = square, and then pad_w is ignored */
    int32_t pad_w;        /* horizontal padding, only read when kw != 0 (vertical padding stays `pad`) */
    int64_t in_x_stride;  /* elements between horizontally adjacent pixels (0 = in_pitch).  A stride smaller than in_c
                             exposes overlapping "wide pixels": the stem runs as a 3x1 conv over 48-channel pixels
                             that each span 3 neighbouring 16-channel space-to-depth cells */
    int64_t in_y_stride;  /* elements between rows   (0 = in_w * in_x_stride) */
    int64_t in_n_stride;  /* elements between images (0 = in_h * in_y_stride) */
    int32_t a_mode;       /* activation fetch: 0 auto, 1 force TMA-im2col, 2 force shifted-patch (stride-1 only) */
    int32_t reserved;     /* flags.  bit 6 (64): the weights are constant -- not written by whatever precedes this launch in the
                             stream -- so the kernel may fetch them before its programmatic-dependency wait (inference programs
                             set it; a training forward that packs weights right before the conv must not).  Tuning / tests:
                             bit 4 (16) row-strided stores instead of the TMA-store epilogue, bit 5 (32) one patch copy per
                             horizontal tap instead of the wide patch; with a forced block_n >= 128 also bit 1 (2) = 256-row
                             tiles, bit 2 (4) = CTA pairs, bits 8.. = cluster size (2|4) for weight-tile multicast */
} y5_conv_desc;

/* Tiling the library will use for a conv: block_k decides the weight packing (cin_pad = ceil(in_c/block_k)*block_k). */
int y5_conv_pick(int32_t in_c, int32_t out_c, int64_t m_rows, int32_t* block_k, int32_t* block_n);

typedef struct y5_conv_plan y5_conv_plan; /* opaque: encoded TMA descriptors + launch geometry for fixed pointers */
int y5_conv_plan_create(const y5_conv_desc* desc, y5_conv_plan** plan);
int y5_conv_plan_run(const y5_conv_plan* plan, void* stream);
void y5_conv_plan_destroy(y5_conv_plan* plan);
/* one-shot convenience: create + run + destroy (tests) */
int y5_conv_bn_silu_fwd(const y5_conv_desc* desc, void* stream);
/* independent direct-convolution kernel (CUDA cores, fp32 accumulate) used by tests to cross-check the tensor-core
 * path on the device; same descriptor, weights in the same packed layout */
int y5_conv_direct_fwd(const y5_conv_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Detect / Segment head level: 1x1 conv with bias + reshape + sigmoid/grid/anchor decode in the GEMM epilogue.
 * Replaces models/yolo.py:95-113 for one level i (conv, view/permute, sigmoid, xy/wh decode, cat into z).
 *   raw : (B, na, ny, nx, no)  un-activated logits, activation dtype            (x[i] of the reference)
 *   z   : (B, z_rows, no) decoded rows; this level writes rows [z_row0, z_row0 + na*ny*nx) of every image
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct y5_detect_desc {
    const void* in;
    int32_t in_pitch;
    int32_t batch, ny, nx, in_c;
    const void* weight;   /* packed [na*no][cin_pad] */
    const float* bias;    /* fp32 [na*no] */
    void* raw;
    void* z;
    int32_t z_rows, z_row0;
    int32_t na, no, nc;   /* no = 5 + nc + nm; columns >= 5+nc (mask coefficients) are passed through un-sigmoided */
    float stride;         /* pixels per cell of this level */
    float anchor_wh[8];   /* na x (w,h) in PIXELS (= anchors * stride), na <= 4 */
    int32_t dtype;
    int32_t block_k;
} y5_detect_desc;
typedef struct y5_detect_plan y5_detect_plan;
int y5_detect_plan_create(const y5_detect_desc* desc, y5_detect_plan** plan);
int y5_detect_plan_run(const y5_detect_plan* plan, void* stream);
/* same plan, outputs redirected to freshly allocated tensors of the same shapes (the reference returns new tensors
 * from every forward; the input side of the plan stays bound to the engine's static buffers) */
int y5_detect_plan_run_to(const y5_detect_plan* plan, void* raw, void* z, void* stream);
void y5_detect_plan_destroy(y5_detect_plan* plan);

/* ---------------------------------------------------------------------------------------------------------------
 * Data-movement kernels (HBM bound)
 * ------------------------------------------------------------------------------------------------------------- */
/* Stem input: NCHW image (Y5_U8 scaled by 1/255, or Y5_F16/Y5_BF16/Y5_F32 already in [0,1]) -> 2x2 space-to-depth
 * NHWC with 16 channels (12 used: (dy*2+dx)*3 + c; 4 zero), so that the 6x6/s2/p2 stem conv of
 * models/yolov5s.yaml:20 becomes a 3x3/s1/p1 conv over 16 channels.  Replaces the `im.half(); im /= 255` of
 * detect.py:206-208 / val.py:259-262 plus the layout change.  h, w even.  `out_row_px` (0 = w/2) is the number of
 * 16-channel cells per output row of the buffer and `out_x_off` the cell where each row starts: the engine keeps one
 * zero cell left and right of every row so the stem conv can read 3 neighbouring cells as one 48-channel pixel. */
int y5_stem_s2d(const void* img, int32_t img_dtype, void* out, int32_t out_dtype, int32_t batch, int32_t h, int32_t w,
                int32_t out_row_px, int32_t out_x_off, void* stream);
/* SPPF pooling (models/common.py:338-340): reads view x (c channels), writes maxpool5, maxpool5^2 (=9x9),
 * maxpool5^3 (=13x13) into three views (usually channel slices 1..3 of the buffer whose slice 0 is x). */
int y5_sppf_pool(const void* x, int32_t x_pitch, void* y1, void* y2, void* y3, int32_t y_pitch, int32_t batch, int32_t h,
                 int32_t w, int32_t c, int32_t ksize, int32_t dtype, void* stream);
/* nn.Upsample(scale_factor=2, mode='nearest') written straight into a (concat) view. */
int y5_upsample2x(const void* x, int32_t x_pitch, void* y, int32_t y_pitch, int32_t batch, int32_t h, int32_t w, int32_t c,
                  int32_t dtype, void* stream);
/* strided channel-slice copy (only needed when a tensor feeds two concat buffers) */
int y5_copy_view(const void* x, int32_t x_pitch, void* y, int32_t y_pitch, int64_t pixels, int32_t c, int32_t dtype,
                 void* stream);
/* NHWC view -> dense NCHW tensor (the API hands NCHW tensors back to PyTorch callers, e.g. Segment's proto) */
int y5_nhwc_to_nchw(const void* x, int32_t x_pitch, void* y, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t dtype,
                    void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * non_max_suppression (utils/general.py:658-767) for a whole batch in one launch set; bit-exact indices.
 *   pred     : (B, N, no) decoded predictions, dtype Y5_F16 | Y5_BF16 | Y5_F32, dense
 *   classes  : optional device array of class ids to keep (int32), n_classes entries
 *   out_rows : (B, max_det, 6+nm) fp32   [x1,y1,x2,y2,conf,cls,masks...]
 *   out_idx  : (B, max_det) int64        candidate id = row*nc + cls of each kept detection
 *   out_count: (B) int32                 detections kept per image
 *   workspace: y5_nms_workspace_bytes(...) bytes of scratch
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct y5_nms_params {
    int32_t batch, n_rows, no, nc, nm;
    int32_t dtype;
    float conf_thres, iou_thres;
    int32_t multi_label, agnostic;
    int32_t max_det;  /* reference default 300 */
    int32_t max_nms;  /* reference constant 30000 */
    float max_wh;     /* reference constant 7680 */
    const int32_t* classes;
    int32_t n_classes;
} y5_nms_params;
int64_t y5_nms_workspace_bytes(const y5_nms_params* p);
int y5_nms_batched(const y5_nms_params* p, const void* pred, float* out_rows, int64_t* out_idx, int32_t* out_count,
                   void* workspace, int64_t workspace_bytes, void* stream);
/* box_iou (ultralytics.utils.metrics.box_iou as used by utils/metrics.py:158,252): (n,4) x (m,4) -> (n,m), fp32 */
int y5_box_iou(const float* a, int32_t n, const float* b, int32_t m, float eps, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * ComputeLoss (utils/loss.py:134-247): build_targets + gather/CIoU/scatter + BCE, forward and backward.
 *   p[l]      : (B, na, ny_l, nx_l, no) logits, dtype Y5_F16 | Y5_BF16 | Y5_F32
 *   targets   : (nt, 6) fp32 [img, cls, x, y, w, h]
 *   anchors   : (nl, na, 2) fp32, grid units
 *   out_loss  : fp32[4] = [loss*bs, lbox, lobj, lcls]  (gains already applied)
 *   grad[l]   : optional, same shape AND dtype as p[l]: d(out_loss[0] * grad_scale)/dp  (fully written)
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct y5_loss_params {
    int32_t nl, batch, na, no, nc;
    int32_t ny[5], nx[5];
    int32_t dtype;
    int32_t nt;
    float anchor_t, box_gain, obj_gain, cls_gain, cls_pw, obj_pw, cp, cn;
    float balance[5];
    float grad_scale;
} y5_loss_params;
int64_t y5_loss_workspace_bytes(const y5_loss_params* p);
/* matches: per level, int32 count + rows (b, a, gj, gi, cls) int64 and tbox fp32 live in the workspace; the
 * build_targets result can be read back with y5_loss_read_targets for parity tests */
int y5_loss_fwd_bwd(const y5_loss_params* p, const void* const* pl, const float* targets, const float* anchors,
                    float* out_loss, void* const* grad, void* workspace, int64_t workspace_bytes, void* stream);
/* same, with the upstream gradient of the loss read from DEVICE memory: grad[l] = d(out_loss[0])/dp * p->grad_scale *
 * (*grad_scale_dev), multiplied in fp32 BEFORE the result is rounded to the prediction dtype -- what autograd does for
 * `scaler.scale(loss).backward()` (train.py:410): a GradScaler factor of 65536 (x WORLD_SIZE) neither overflows fp16 on
 * the way nor flushes small objectness gradients to zero.  grad_scale_dev may be NULL (= 1).  Targets whose image index
 * is outside [0, batch) or whose class is outside [0, nc) are ignored (the reference raises IndexError for them). */
int y5_loss_fwd_bwd_scaled(const y5_loss_params* p, const void* const* pl, const float* targets, const float* anchors,
                           float* out_loss, void* const* grad, const float* grad_scale_dev, void* workspace,
                           int64_t workspace_bytes, void* stream);
int y5_loss_read_targets(const y5_loss_params* p, const void* workspace, int32_t level, int64_t* idx5_host,
                         float* tbox_host, int32_t* count_host, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training-mode Conv = SiLU(BN(conv(x)))  (reference models/common.py:86-88; gradients of the same).
 * Forward:   y = conv(x, W)            -> y5_conv_bn_silu_fwd with act = Y5_ACT_NONE and a zero bias (same kernel)
 *            column sums of y          -> y5_bn_stats
 *            z = act(bn(y))            -> y5_bn_act_fwd    (derives mean/invstd, updates running_mean / running_var)
 * Backward:  dy, dgamma, dbeta         -> y5_bn_act_bwd
 *            dW                        -> y5_conv_wgrad    (tcgen05, MN-major operands straight from NHWC)
 *            dx = conv(dy, W^T flipped)-> y5_conv_bn_silu_fwd again on transposed/flipped packed weights (stride-2
 *                                         layers first expand dy with y5_zero_stuff2x); `residual` = dx accumulates.
 * Detect.m[i] (models/yolo.py:97) has a bias and no BN: its bias gradient is y5_col_sum(dy).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct y5_wgrad_desc {
    const void* in;       /* x view, NHWC: element (n,y,x,c) at ((n*in_h + y)*in_w + x)*in_pitch + c */
    int32_t in_pitch;
    int32_t batch, in_h, in_w, in_c;
    const void* dout;     /* dy view [batch*Ho*Wo][out_c], row pitch dout_pitch elements */
    int32_t dout_pitch;
    int32_t out_c;
    float* dweight;       /* fp32 [out_c][ksize][ksize][in_c] (KRSC); zeroed by the call unless accumulate != 0 */
    int32_t ksize, stride, pad;
    int32_t dtype;        /* Y5_F16 | Y5_BF16 (x and dy) */
    int32_t accumulate;
    int32_t reserved;
    /* optional, as in y5_conv_desc: kw != 0 -> filter is ksize x kw with horizontal padding pad_w (dweight is
     * [out_c][ksize][kw][in_c]); non-zero strides (elements) describe a general NHWC view, e.g. overlapping "wide pixels" */
    int32_t kw, pad_w;
    int64_t in_x_stride, in_y_stride, in_n_stride;
} y5_wgrad_desc;
int y5_conv_wgrad(const y5_wgrad_desc* d, void* stream);

/* workspace of y5_bn_stats / y5_bn_act_bwd / y5_col_sum: 2 * channels doubles.  For y5_bn_stats and y5_bn_act_bwd it
 * must be ZERO on entry and is left dirty (callers carve it from one arena cleared once per step); y5_col_sum clears
 * its own. */
int64_t y5_bn_workspace_bytes(int32_t channels);
/* per-channel sum and sum of squares of a [rows][channels] view, accumulated into workspace (fp64) */
int y5_bn_stats(const void* y, int32_t pitch, int64_t rows, int32_t channels, int32_t dtype, void* workspace,
                void* stream);
/* z = act(gamma * (y - mean) * invstd + beta), act: Y5_ACT_NONE | Y5_ACT_SILU; z may be a channel-slice view.
 * sums != NULL (training): mean / invstd (biased variance + eps) are first derived from the y5_bn_stats workspace and
 * WRITTEN to mean / invstd, and running_mean / running_var (nullable) are updated like nn.BatchNorm2d does (momentum,
 * unbiased variance).  sums == NULL (eval): mean / invstd are inputs.  residual != NULL adds a view of the same shape
 * after the activation (Bottleneck shortcut, models/common.py:181); its gradient is dz itself. */
int y5_bn_act_fwd(const void* y, int32_t y_pitch, void* z, int32_t z_pitch, int64_t rows, int32_t channels,
                  int32_t dtype, float* mean, float* invstd, const float* gamma, const float* beta, int32_t act,
                  const void* sums, float eps, float momentum, float* running_mean, float* running_var,
                  const void* residual, int32_t res_pitch, void* stream);
/* given dz: dy (gradient w.r.t. the conv output), dgamma, dbeta (fp32, overwritten).  Two launches: the reduce pass forms
 * du = dz * act'(bn(y)) and its two column sums and, for Y5_ACT_SILU, parks du in the dy buffer; the apply pass turns it into dy
 * in place.  dy may alias dz (every element is read before it is written) but not y. */
int y5_bn_act_bwd(const void* y, int32_t y_pitch, const void* dz, int32_t dz_pitch, void* dy, int32_t dy_pitch,
                  int64_t rows, int32_t channels, int32_t dtype, const float* mean, const float* invstd,
                  const float* gamma, const float* beta, int32_t act, float* dgamma, float* dbeta, void* workspace,
                  void* stream);
/* out[c] = sum over rows of y[row][c] (fp32) */
int y5_col_sum(const void* y, int32_t pitch, int64_t rows, int32_t channels, int32_t dtype, float* out, void* workspace,
               void* stream);
/* OIHW master weights (Y5_F32 | Y5_F16 | Y5_BF16) -> K-major packings in `dtype` for y5_conv_bn_silu_fwd, either may be
 * NULL:  fwd [out_c][k][k][in_c_pad] = w[o][i][r][s]  and  dgrad [in_c][k][k][out_c_pad] = w[o][i][k-1-r][k-1-s]
 * (padding channels zero; pads = channel counts rounded up to the block_k y5_conv_pick returns). */
int y5_weight_pack(const void* w, int32_t w_dtype, int32_t out_c, int32_t in_c, int32_t ksize, void* fwd, int32_t in_c_pad,
                   void* dgrad, int32_t out_c_pad, int32_t dtype, void* stream);
/* The same for every filter of a model in ONE launch (the training step re-packs its fp32 master weights once per forward):
 * `items` (device memory) describes the filters exactly like the arguments of y5_weight_pack; block b of the launch converts
 * elements [chunk_index[b] * y5_weight_pack_chunk_elems(), ...) of the concatenated [fwd | dgrad] packing of item chunk_item[b]
 * (both arrays in device memory, built once per model by the caller). */
typedef struct y5_pack_item {
    const void* w;
    void* fwd;
    void* dgrad;
    int32_t w_dtype;
    int32_t out_c, in_c, ksize;
    int32_t in_c_pad, out_c_pad;
} y5_pack_item;
int32_t y5_weight_pack_chunk_elems(void);
int y5_weight_pack_multi(const y5_pack_item* items, const int32_t* chunk_item, const int32_t* chunk_index, int32_t n_chunks,
                         int32_t dtype, void* stream);
/* backward of y5_upsample2x: dx[n,i,j,:] = dy[n,2i,2j,:] + dy[n,2i,2j+1,:] + dy[n,2i+1,2j,:] + dy[n,2i+1,2j+1,:] */
int y5_upsample2x_bwd(const void* dy, int32_t dy_pitch, void* dx, int32_t dx_pitch, int32_t batch, int32_t h, int32_t w,
                      int32_t c, int32_t dtype, void* stream);
/* backward of y5_sppf_pool + the concat of SPPF (models/common.py:338-340): cat = [a, m(a), m(m(a)), m(m(m(a)))] as 4
 * channel slices of c channels (the forward's buffer), dcat its gradient; writes da.  Arg-max ties resolve to the
 * first maximum in row-major window order, as torch's max_pool2d backward does.  workspace: 3*B*h*w*c floats. */
int64_t y5_sppf_bwd_workspace_bytes(int32_t batch, int32_t h, int32_t w, int32_t c);
int y5_sppf_pool_bwd(const void* cat, int32_t cat_pitch, const void* dcat, int32_t dcat_pitch, void* da, int32_t da_pitch,
                     int32_t batch, int32_t h, int32_t w, int32_t c, int32_t ksize, int32_t dtype, void* workspace,
                     void* stream);
/* y[n, 2i, 2j, :] = x[n, i, j, :], other pixels of the (2h, 2w) output zero */
int y5_zero_stuff2x(const void* x, int32_t x_pitch, void* y, int32_t y_pitch, int32_t batch, int32_t h, int32_t w,
                    int32_t c, int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * The callers either side of the hot path, batched on the device, no host synchronisation.
 * ------------------------------------------------------------------------------------------------------------------ */
/* Pre-processing: letterbox (utils/augmentations.py:85-115: cv2.resize INTER_LINEAR to (new_w,new_h), constant border) +
 * BGR->RGB + HWC->CHW (utils/dataloaders.py:354-357) + uint8 -> float /255 (detect.py:205-208, val.py:259-262,
 * models/common.py:924-926) for a batch of device-resident uint8 HWC images of different sizes.  The resize is OpenCV's
 * fixed-point bilinear kernel restated bit for bit (oracle/pre_ref.py, pinned against the installed cv2).  The geometry
 * (new_w/new_h/top/left) is computed by the caller exactly as the reference's letterbox() does. */
typedef struct y5_letterbox_image {
    const void* data;       /* uint8 HWC, 3 channels, device memory */
    int32_t src_h, src_w;
    int32_t row_bytes;      /* bytes between source rows (>= 3*src_w) */
    int32_t new_h, new_w;   /* size after cv2.resize ("new_unpad") */
    int32_t top, left;      /* border offsets inside the (out_h, out_w) canvas */
} y5_letterbox_image;
int y5_letterbox_max_images(void); /* descriptors are consumed in groups of this many per launch */
/* out: out_dtype Y5_U8 -> (n,3,out_h,out_w) bytes (what the dataloader yields); Y5_F16/BF16/F32 -> the same /255;
 * s2d != 0 (fp16/bf16 only) -> the stem's 2x2 space-to-depth cells, layout and out_row_px/out_x_off as in y5_stem_s2d.
 * `images` is a HOST array (copied into the launch parameters). */
int y5_letterbox(const y5_letterbox_image* images, int32_t n_images, int32_t out_h, int32_t out_w, int32_t swap_rb,
                 int32_t pad_value, void* out, int32_t out_dtype, int32_t s2d, int32_t out_row_px, int32_t out_x_off,
                 void* stream);

/* process_mask (utils/segment/general.py:25-52, crop_mask :10-22): for detection i of image img_index[i] (NULL = image 0):
 * sigmoid(coef_i . protos[img]) at mask resolution, zeroed outside the box scaled by (mw/in_w, mh/in_h), optionally
 * bilinearly up-sampled (align_corners=False) to (in_h,in_w), thresholded at 0.5.
 *   mode 0: process_mask(upsample=False) -> (n, mh, mw);  mode 1: process_mask(upsample=True) -> (n, in_h, in_w);
 *   mode 2: process_mask_native (:55-76): no low-resolution crop, the prototype window [top, left, height, width] (`window`,
 *           HOST array of 4 ints) is up-sampled to (in_h, in_w) and cropped to the un-scaled boxes -> (n, in_h, in_w)
 *   protos (batch, c, mh, mw) NCHW Y5_F16|Y5_BF16|Y5_F32 (read as float, like `protos.float()`); coef rows of `coef_stride`
 *   floats (may point at column 6 of the NMS rows); boxes rows of `box_stride` floats, xyxy in network-input pixels;
 *   out Y5_F32 {0,1} (the reference's `masks.gt_(0.5)`) or Y5_U8.
 * Detections of the same image must be adjacent.  workspace: y5_process_mask_workspace_bytes (modes 1, 2). */
int64_t y5_process_mask_workspace_bytes(int32_t n, int32_t mh, int32_t mw, int32_t mode);
int y5_process_mask(const void* protos, int32_t proto_dtype, int32_t batch, int32_t c, int32_t mh, int32_t mw, const float* coef,
                    int32_t coef_stride, const float* boxes, int32_t box_stride, const int32_t* img_index, int32_t n,
                    int32_t in_h, int32_t in_w, int32_t mode, const int32_t* window, void* out, int32_t out_dtype,
                    void* workspace, int64_t workspace_bytes, void* stream);
/* crop_mask (utils/segment/general.py:10-22): out = masks * [box contains the pixel]; masks/out (n,h,w) fp32 */
int y5_crop_mask(const float* masks, const float* boxes, int32_t box_stride, int32_t n, int32_t h, int32_t w, float* out,
                 void* stream);

/* scale_boxes + clip_boxes (utils/general.py:613-626), in place on rows of `row_stride` floats (xyxy first).
 * meta: per image 5 floats [gain, pad_x, pad_y, w0, h0] (device).  Image of row i = img_index[i], or i / rows_per_image
 * when img_index is NULL (then `count`, if given, limits each image to its first count[img] rows: the padded layout
 * y5_nms_batched produces), or 0 when both are absent. */
int y5_scale_boxes(float* boxes, int32_t row_stride, int64_t n_rows, const int32_t* img_index, int32_t rows_per_image,
                   const int32_t* count, const float* meta, void* stream);
/* val.py:303-306 for the whole batch: targets (nt,6) [img, cls, cx, cy, w, h] in network-input pixels ->
 * out (nt,6) [img, cls, x1, y1, x2, y2] in native pixels (xywh2xyxy, then scale_boxes with the image's meta). */
int y5_labels_native(const float* targets, int32_t nt, const float* meta, float* out, void* stream);
/* process_batch (utils/metrics.py:224-265, box branch) for every image of a batch in one launch:
 *   det (batch, max_det, row_stride>=6) fp32 rows [x1,y1,x2,y2,conf,cls,...] in native pixels (image b at det + b*img_stride),
 *   count[b] valid rows (NULL = max_det); labels (nt,6) [img, cls, x1,y1,x2,y2] in any order; iouv (niou) thresholds;
 *   correct (batch, max_det, niou) uint8: the reference's boolean matrix (rows >= count[b] are 0).  Bit-exact. */
int y5_match_batch(const float* det, int64_t img_stride, int32_t row_stride, const int32_t* count, int32_t batch,
                   int32_t max_det, const float* labels, int32_t nt, const float* iouv, int32_t niou, float eps,
                   uint8_t* correct, void* stream);

/* Fused optimizer step (train.py:413-421): un-scale + clip_grad_norm_ + SGD(momentum, nesterov) over parameter groups
 * (utils/torch_utils.py:256-289) + optimizer.zero_grad + ModelEMA.update (utils/torch_utils.py:359-368) in two
 * multi-tensor launches.  All tensors fp32.  `table`, `chunk_*`, `hyper`, `partial` are DEVICE arrays owned by the caller:
 *   table[t]            one entry per tensor (buffers take part in the EMA only: grad = mom = NULL)
 *   chunk_tensor/index  block c handles elements [chunk_index[c]*y5_opt_chunk_elems(), +y5_opt_chunk_elems()) of tensor chunk_tensor[c]
 *   hyper               floats: [Y5_OPT_INV_SCALE] 1/loss scale, [Y5_OPT_MAX_NORM] clip norm (<= 0: off), [Y5_OPT_EMA_DECAY],
 *                       [Y5_OPT_EMA_TAU], [Y5_OPT_EMA_UPDATES] counter (advanced by the call when do_ema),
 *                       [Y5_OPT_OUT_NORM] total gradient norm (written), [Y5_OPT_OUT_SKIPPED] 1 if a non-finite gradient
 *                       made the step skip (written), then per group g at Y5_OPT_GROUPS + 4g: lr, momentum, weight_decay, nesterov
 *   partial             2 * n_chunks floats of scratch */
typedef struct y5_opt_tensor {
    void* param;
    void* grad;
    void* mom;
    void* ema;
    int64_t numel;
    int32_t group;
    int32_t reserved;
} y5_opt_tensor;
#define Y5_OPT_INV_SCALE 0
#define Y5_OPT_MAX_NORM 1
#define Y5_OPT_EMA_DECAY 2
#define Y5_OPT_EMA_TAU 3
#define Y5_OPT_EMA_UPDATES 4
#define Y5_OPT_OUT_NORM 5
#define Y5_OPT_OUT_SKIPPED 6
#define Y5_OPT_GROUPS 8
int32_t y5_opt_chunk_elems(void);
int y5_opt_step(const y5_opt_tensor* table, const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t n_chunks,
                float* hyper, float* partial, int32_t do_step, int32_t do_ema, int32_t zero_grad, void* stream);

/* Data-parallel gradient exchange, device half (utils/torch_utils.py:61-70 smart_DDP / train.py:404-414): copy every gradient
 * of `table` (entries with mom != NULL; a NULL grad contributes zeros) into ONE contiguous fp32 arena in a single launch, so the
 * all-reduce is one NCCL call over the arena and y5_opt_step reads the averaged gradients from it (a second table whose grad
 * pointers are arena + arena_offset[t]).  arena_offset[t]: element offset of tensor t, a multiple of 4; arena 16-byte aligned.
 * present[t] = 1.0 / 0.0: tensor t had a gradient on this rank (callers place `present` right behind the arena so the same
 * all-reduce averages it); y5_grad_bind then rewrites the arena table's gradient pointers -- NULL where present[t] == 0 on every
 * rank -- so the update skips such parameters like the single-process step does. */
int y5_grad_pack(const y5_opt_tensor* table, const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t n_chunks,
                 const int64_t* arena_offset, float* arena, float* present, void* stream);
int y5_grad_bind(y5_opt_tensor* arena_table, int32_t n_tensors, const int64_t* arena_offset, float* arena, const float* present,
                 void* stream);

/* Fold eval-mode BatchNorm into a conv and pack it for y5_conv_bn_silu_fwd in ONE launch (utils/torch_utils.py:224-254):
 *   packed[o][r][s][i_pad] = w[o][i][r][s] * gamma[o] / sqrt(var[o] + eps)   (activation dtype, zero padded)
 *   bias_out[o]            = beta[o] + (conv_bias[o] - mean[o]) * gamma[o] / sqrt(var[o] + eps)   (fp32)
 * gamma == NULL: no BatchNorm (already fused conv): packed = w, bias_out = conv_bias (or 0).  w: OIHW Y5_F32|F16|BF16;
 * BN tensors fp32 or the weight dtype (`bn_dtype`).  out_c_pad >= out_c rows are written (extra rows zero, bias 0). */
int y5_fold_pack(const void* w, int32_t w_dtype, int32_t out_c, int32_t in_c, int32_t kh, int32_t kw, const void* conv_bias,
                 const void* gamma, const void* beta, const void* mean, const void* var, int32_t bn_dtype, float eps,
                 void* packed, int32_t in_c_pad, int32_t out_c_pad, float* bias_out, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* Y5B200_H */
