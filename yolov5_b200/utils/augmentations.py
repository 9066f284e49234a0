"""The step right before the hot path, on the device: letterbox resize/pad + BGR->RGB + HWC->CHW + /255
(reference utils/augmentations.py:85-115, utils/dataloaders.py:354-357, detect.py:205-208, models/common.py:924-926).

``letterbox`` keeps the reference's signature and return value ``(image, ratio, (dw, dh))``; the resize itself runs in
liby5b200 (y5_letterbox: OpenCV's fixed-point bilinear kernel restated bit for bit, so the letterboxed bytes equal what
``cv2.resize`` + ``cv2.copyMakeBorder`` produce).  ``letterbox_batch`` is the batched form the engine wants: a list of
device-resident uint8 HWC images of different sizes -> one (B,3,H,W) tensor (uint8, or already scaled to [0,1] in
fp16/bf16/fp32), one launch per 24 images, no host synchronisation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib


def letterbox_geometry(shape_hw, new_shape=(640, 640), auto=True, scaleFill=False, scaleup=True, stride=32):
    """(new_unpad (w, h), ratio (w, h), (dw, dh), (top, bottom, left, right)) as utils/augmentations.py:85-113 derives them."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    h, w = int(shape_hw[0]), int(shape_hw[1])
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:  # only scale down (better val mAP)
        r = min(r, 1.0)
    ratio = r, r
    new_unpad = round(w * r), round(h * r)
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:  # minimum rectangle
        dw, dh = float(np.mod(dw, stride)), float(np.mod(dh, stride))
    elif scaleFill:  # stretch
        dw, dh = 0.0, 0.0
        new_unpad = (new_shape[1], new_shape[0])
        ratio = new_shape[1] / w, new_shape[0] / h
    dw /= 2
    dh /= 2
    top, bottom = round(dh - 0.1), round(dh + 0.1)
    left, right = round(dw - 0.1), round(dw + 0.1)
    return new_unpad, ratio, (dw, dh), (top, bottom, left, right)


def _as_device_hwc(im, device):
    if isinstance(im, np.ndarray):
        im = torch.from_numpy(np.ascontiguousarray(im))
    if not isinstance(im, torch.Tensor) or im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
        raise TypeError("y5b200: letterbox expects uint8 HWC images with 3 channels")
    if not im.is_cuda:
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("y5b200: letterbox runs in liby5b200 on a CUDA device (no CPU / OpenCV fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        im = im.to(device, non_blocking=True)
    if im.stride(2) != 1 or im.stride(1) != 3:
        im = im.contiguous()
    return im


def letterbox_batch(images, new_shape=(640, 640), color=(114, 114, 114), auto=False, scaleFill=False, scaleup=True, stride=32,
                    swap_rb=True, dtype=torch.uint8, device=None, out=None, s2d_out=None):
    """images: list of uint8 HWC (BGR unless swap_rb=False) tensors / arrays of any sizes.  All images are letterboxed to
    the SAME canvas: `new_shape` (auto=False), or the minimum-rectangle canvas of the first image (auto=True, as the
    reference does per image: then every image must produce that canvas).  Returns (batch, ratios, pads):
    batch (B,3,H,W) `dtype` (uint8 = the dataloader's bytes; float dtypes are divided by 255), CHW with RGB order when
    swap_rb.  `s2d_out` = (buffer, row_px, x_off) writes the stem's space-to-depth cells instead (engine internal)."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    if len(set(color)) != 1:
        raise NotImplementedError("y5b200: letterbox border colour must be grey (the reference uses (114, 114, 114))")
    ims = [_as_device_hwc(im, device) for im in images]
    dev = ims[0].device
    geo = [letterbox_geometry(im.shape[:2], new_shape, auto, scaleFill, scaleup, stride) for im in ims]
    canvases = {(g[0][1] + g[3][0] + g[3][1], g[0][0] + g[3][2] + g[3][3]) for g in geo}
    if len(canvases) != 1:
        raise ValueError(f"y5b200: letterbox_batch needs one output shape for the whole batch, got {sorted(canvases)} (use auto=False)")
    out_h, out_w = canvases.pop()
    lib = _lib.lib()
    arr = (_lib.LetterboxImage * len(ims))()
    for d, im, g in zip(arr, ims, geo):
        d.data, d.src_h, d.src_w, d.row_bytes = im.data_ptr(), im.shape[0], im.shape[1], im.stride(0)
        d.new_w, d.new_h = g[0]
        d.top, d.left = g[3][0], g[3][2]
    with _lib.on(dev):
        if s2d_out is not None:
            buf, row_px, x_off = s2d_out
            _lib.check(lib.y5_letterbox(arr, len(ims), out_h, out_w, int(swap_rb), int(color[0]), buf.data_ptr(), _lib.dtype_code(buf.dtype), 1,
                                        row_px, x_off, C.c_void_p(_lib.stream_ptr(dev))), "letterbox")
            res = buf
        else:
            res = out if out is not None else torch.empty(len(ims), 3, out_h, out_w, dtype=dtype, device=dev)
            assert res.shape == (len(ims), 3, out_h, out_w) and res.is_contiguous()
            _lib.check(lib.y5_letterbox(arr, len(ims), out_h, out_w, int(swap_rb), int(color[0]), res.data_ptr(), _lib.dtype_code(res.dtype), 0, 0,
                                        0, C.c_void_p(_lib.stream_ptr(dev))), "letterbox")
    for im in ims:  # the launch reads these buffers asynchronously: tie their lifetime to the stream
        im.record_stream(torch.cuda.current_stream(dev))
    return res, [g[1] for g in geo], [g[2] for g in geo]


def letterbox(im, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleFill=False, scaleup=True, stride=32):
    """Reference signature (utils/augmentations.py:85): one HWC uint8 image -> (letterboxed HWC image, ratio, (dw, dh)).
    numpy in -> numpy out (the bytes cv2 would produce), CUDA tensor in -> CUDA tensor out."""
    is_np = isinstance(im, np.ndarray)
    batch, ratios, pads = letterbox_batch([im], new_shape, color, auto, scaleFill, scaleup, stride, swap_rb=False)
    hwc = batch[0].permute(1, 2, 0).contiguous()
    return (hwc.cpu().numpy() if is_np else hwc), ratios[0], pads[0]
