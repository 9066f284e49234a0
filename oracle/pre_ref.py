"""ORACLE (test infrastructure, never on the product path): CPU restatement of the step right BEFORE the hot path --
letterbox resize/pad + HWC->CHW + BGR->RGB + /255.  numpy integer arithmetic.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this.

Reference lines restated (paths relative to /root/reference):
  utils/augmentations.py:85-115   letterbox (geometry: ratio, new_unpad, dw/dh split, border 114)   -> letterbox_geometry, letterbox
  utils/dataloaders.py:354-357    im.transpose((2,0,1))[::-1] (HWC BGR -> CHW RGB)                   -> to_chw_rgb
  detect.py:205-208, val.py:259-262  uint8 -> fp16/fp32, /255                                        -> to_chw_rgb(normalise=True)
Third-party arithmetic on this step: OpenCV ``cv2.resize(..., INTER_LINEAR)`` on uint8 (opencv-python 4.13.0 is installed
in this image; its source is not under /root/reference).  `resize_linear_u8` restates OpenCV's fixed-point bilinear
kernel (resize.cpp: coefficients scaled by 2^11 and saturate-cast to int16, horizontal pass into int32 rows, vertical
pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2) and is PINNED bit-exactly against the installed cv2 by
tests/golden/make_golden.py (gen_pre) and, when cv2 is importable, by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 11
INTER_ONE = 1 << INTER_BITS


def _axis_coeffs(src: int, dst: int, horizontal: bool):
    """Per destination index: the two source indices and the two int16 weights, as OpenCV's resize.cpp computes them.
    Horizontally an out-of-range tap collapses onto the border pixel with weight (2048, 0); vertically the two row
    indices are clamped individually and the fractional weights are KEPT (so border rows are a two-term sum of the same
    source row, which the fixed-point rounding makes differ from the one-term form by up to 1)."""
    scale = 1.0 / (np.float64(dst) / np.float64(src))  # scale_x = 1. / inv_scale_x
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)  # fx = (float)((dx + 0.5) * scale_x - 0.5)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if horizontal:
        lo = s < 0
        f[lo] = 0.0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0.0
        s[hi] = src - 1
    w1 = np.rint(f.astype(np.float32) * np.float32(INTER_ONE)).astype(np.int64)  # saturate_cast<short>(float): round half to even
    w0 = np.rint((np.float32(1.0) - f) * np.float32(INTER_ONE)).astype(np.int64)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), w0, w1


def resize_linear_u8(img: np.ndarray, dst_wh) -> np.ndarray:
    """cv2.resize(img, dst_wh, interpolation=cv2.INTER_LINEAR) for uint8 HWC images, bit-exact."""
    h, w = img.shape[:2]
    dw, dh = int(dst_wh[0]), int(dst_wh[1])
    x0, x1, a0, a1 = _axis_coeffs(w, dw, True)
    y0, y1, b0, b1 = _axis_coeffs(h, dh, False)
    src = img.astype(np.int64)
    rows = src[:, x0] * a0[None, :, None] + src[:, x1] * a1[None, :, None]  # (h, dw, c) scaled by 2^11
    s0, s1 = rows[y0], rows[y1]
    out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_geometry(shape_hw, new_shape=(640, 640), auto=True, scaleFill=False, scaleup=True, stride=32):
    """(new_unpad (w,h), ratio (w,h), (dw, dh) halves, (top, bottom, left, right)) exactly as utils/augmentations.py:85-113
    derives them (python round = banker's rounding, np.mod for the minimum-rectangle padding)."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    h, w = int(shape_hw[0]), int(shape_hw[1])
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    ratio = (r, r)
    new_unpad = (round(w * r), round(h * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = float(np.mod(dw, stride)), float(np.mod(dh, stride))
    elif scaleFill:
        dw, dh = 0.0, 0.0
        new_unpad = (new_shape[1], new_shape[0])
        ratio = (new_shape[1] / w, new_shape[0] / h)
    dw /= 2
    dh /= 2
    top, bottom = round(dh - 0.1), round(dh + 0.1)
    left, right = round(dw - 0.1), round(dw + 0.1)
    return new_unpad, ratio, (dw, dh), (top, bottom, left, right)


def letterbox(im: np.ndarray, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleFill=False, scaleup=True, stride=32):
    """Returns (padded HWC uint8 image, ratio, (dw, dh)) like the reference."""
    new_unpad, ratio, (dw, dh), (top, bottom, left, right) = letterbox_geometry(im.shape[:2], new_shape, auto, scaleFill, scaleup, stride)
    if (im.shape[1], im.shape[0]) != tuple(new_unpad):
        im = resize_linear_u8(im, new_unpad)
    out = np.empty((im.shape[0] + top + bottom, im.shape[1] + left + right, im.shape[2]), np.uint8)
    out[...] = np.asarray(color, np.uint8)[None, None, :]
    out[top : top + im.shape[0], left : left + im.shape[1]] = im
    return out, ratio, (dw, dh)


def to_chw_rgb(im_hwc_bgr: np.ndarray, normalise: bool = False) -> np.ndarray:
    """HWC BGR -> CHW RGB (utils/dataloaders.py:356); with `normalise`, float32 / 255 (detect.py:206-208)."""
    x = np.ascontiguousarray(im_hwc_bgr.transpose(2, 0, 1)[::-1])
    return x.astype(np.float32) / np.float32(255) if normalise else x


def synth_image(h: int, w: int, seed: int) -> np.ndarray:
    """Seeded uint8 HWC image with smooth structure plus noise (so interpolation weights matter)."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = np.stack([127 + 120 * np.sin(xx / (7 + c) + yy / (11 - c) + c) for c in range(3)], -1)
    return np.clip(base + rs.uniform(-30, 30, (h, w, 3)), 0, 255).astype(np.uint8)
