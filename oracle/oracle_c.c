/* ORACLE (test infrastructure): plain-C restatement of the greedy suppression and pairwise IoU used on the
 * NMS / metric path.  Never linked into the product library.
 *
 *  y5o_nms_greedy  -- torchvision.ops.nms semantics as called from reference utils/general.py:750
 *                     (torchvision 0.26 CPU kernel; source not vendored in the reference -- behaviour pinned
 *                     against the installed op by tests/golden/make_golden.py):
 *                       boxes already sorted by score descending (stable); area = (x2-x1)*(y2-y1);
 *                       suppress j>i when inter/(area_i+area_j-inter) > thr (strict, NaN never suppresses);
 *                       the fp32 ratio is compared against the threshold as a DOUBLE (the CPU kernel's
 *                       signature is nms_kernel_impl(dets, scores, double iou_threshold)).  torchvision's CUDA
 *                       kernel narrows the threshold to float instead; the two differ only when the ratio falls
 *                       between float(thr) and thr.  We follow the CPU kernel: it is the one that can be run
 *                       (and therefore pinned) where the fixtures are generated.
 *  y5o_box_iou     -- ultralytics.utils.metrics.box_iou as used by reference utils/metrics.py:158,252:
 *                       inter/(area1+area2-inter+eps), fp32.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile); no FMA contraction so that every
 * operation rounds to fp32 exactly like the reference's scalar code.
 */
#include <stdint.h>
#include <stdlib.h>

static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }

/* boxes: n x 4 (x1,y1,x2,y2) in score order. keep: out indices (capacity n). returns number kept (<= max_keep). */
int64_t y5o_nms_greedy(const float* boxes, int64_t n, double thr, int64_t max_keep, int64_t* keep) {
    unsigned char* dead = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
    int64_t nk = 0;
    for (int64_t i = 0; i < n && nk < max_keep; ++i) {
        if (dead[i]) continue;
        keep[nk++] = i;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        const float iarea = (ix2 - ix1) * (iy2 - iy1);
        for (int64_t j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            const float xx1 = fmaxf_(ix1, boxes[4 * j]), yy1 = fmaxf_(iy1, boxes[4 * j + 1]);
            const float xx2 = fminf_(ix2, boxes[4 * j + 2]), yy2 = fminf_(iy2, boxes[4 * j + 3]);
            const float w = fmaxf_(0.0f, xx2 - xx1), h = fmaxf_(0.0f, yy2 - yy1);
            const float inter = w * h;
            const float jarea = (boxes[4 * j + 2] - boxes[4 * j]) * (boxes[4 * j + 3] - boxes[4 * j + 1]);
            const float ovr = inter / (iarea + jarea - inter);
            if ((double)ovr > thr) dead[j] = 1; /* torchvision CPU kernel: iou_threshold is a double */
        }
    }
    free(dead);
    return nk;
}

/* a: n x 4, b: m x 4, out: n x m */
void y5o_box_iou(const float* a, int64_t n, const float* b, int64_t m, float eps, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        const float area1 = (a[4 * i + 2] - a[4 * i]) * (a[4 * i + 3] - a[4 * i + 1]);
        for (int64_t j = 0; j < m; ++j) {
            const float area2 = (b[4 * j + 2] - b[4 * j]) * (b[4 * j + 3] - b[4 * j + 1]);
            float w = fminf_(a[4 * i + 2], b[4 * j + 2]) - fmaxf_(a[4 * i], b[4 * j]);
            float h = fminf_(a[4 * i + 3], b[4 * j + 3]) - fmaxf_(a[4 * i + 1], b[4 * j + 1]);
            w = w < 0.0f ? 0.0f : w;
            h = h < 0.0f ? 0.0f : h;
            const float inter = w * h;
            out[i * m + j] = inter / (area1 + area2 - inter + eps);
        }
    }
}
