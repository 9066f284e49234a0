"""Mask post-processing of the segmentation path on the device (reference utils/segment/general.py:10-76): the step
right after non_max_suppression for SegmentationModel outputs.  Same signatures as the reference; the arithmetic is
liby5b200's y5_process_mask (coefficients x prototypes + sigmoid + crop + bilinear up-sampling + threshold in two launches,
no intermediate (n,h,w) tensors in PyTorch)."""
from __future__ import annotations

import ctypes as C

import torch

from ... import _lib


def _check_cuda(*ts):
    if not all(isinstance(t, torch.Tensor) and t.is_cuda for t in ts):
        raise RuntimeError("y5b200: mask post-processing runs on CUDA tensors only (no CPU / PyTorch fallback)")


def crop_mask(masks, boxes):
    """Zero everything outside each mask's xyxy box (reference :10-22).  masks (n,h,w), boxes (n,4) in mask pixels."""
    _check_cuda(masks, boxes)
    m = masks.float().contiguous()
    b = boxes.float().contiguous()
    out = torch.empty_like(m)
    n, h, w = m.shape
    if n:
        with _lib.on(m.device):
            _lib.check(_lib.lib().y5_crop_mask(m.data_ptr(), b.data_ptr(), b.stride(0), n, h, w, out.data_ptr(),
                                               C.c_void_p(_lib.stream_ptr(m.device))), "crop_mask")
    return out.to(masks.dtype)


def process_mask_batch(protos, coef, boxes, img_index, shape, upsample=False, out_dtype=torch.float32, native=False):
    """Batched form: protos (B,c,mh,mw) as the model returns them; coef (n,c) / boxes (n,4) fp32 rows of the kept
    detections of ALL images (image-major order, they may be strided views into the NMS rows), img_index (n,) int32 or
    None (single image).  Returns (n, mh, mw) or (n, ih, iw) masks of {0,1} in `out_dtype` (float32 like the reference's
    `masks.gt_(0.5)`, or uint8).  native=True is process_mask_native's order of operations (up-sample, then crop)."""
    _check_cuda(protos, coef, boxes)
    if protos.dim() == 3:
        protos = protos[None]
    b, c, mh, mw = protos.shape
    ih, iw = int(shape[0]), int(shape[1])
    n = coef.shape[0]
    dev = protos.device
    mode = 2 if native else (1 if upsample else 0)
    oh, ow = (ih, iw) if mode else (mh, mw)
    out = torch.empty(n, oh, ow, dtype=out_dtype, device=dev)
    if n == 0:
        return out
    protos = protos.contiguous()
    if coef.dtype != torch.float32 or coef.stride(1) != 1:
        coef = coef.float().contiguous()
    if boxes.dtype != torch.float32 or boxes.stride(1) != 1:
        boxes = boxes.float().contiguous()
    idx = None
    if img_index is not None:
        idx = img_index.to(dev, torch.int32).contiguous()
    window = None
    if native:  # reference :68-72: the un-padded window of the prototype map
        gain = min(mh / ih, mw / iw)
        pad = (mw - iw * gain) / 2, (mh - ih * gain) / 2
        top, left = int(pad[1]), int(pad[0])
        bottom, right = int(mh - pad[1]), int(mw - pad[0])
        window = (C.c_int32 * 4)(top, left, bottom - top, right - left)
    lib = _lib.lib()
    need = int(lib.y5_process_mask_workspace_bytes(n, mh, mw, mode))
    ws = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    with _lib.on(dev):
        _lib.check(lib.y5_process_mask(protos.data_ptr(), _lib.dtype_code(protos.dtype), b, c, mh, mw, coef.data_ptr(), coef.stride(0),
                                       boxes.data_ptr(), boxes.stride(0), idx.data_ptr() if idx is not None else None, n, ih, iw, mode, window,
                                       out.data_ptr(), _lib.dtype_code(out_dtype), ws.data_ptr(), need, C.c_void_p(_lib.stream_ptr(dev))),
                   "process_mask")
    return out


def process_mask(protos, masks_in, bboxes, shape, upsample=False):
    """Reference signature (utils/segment/general.py:25-52): protos (c,mh,mw), masks_in (n,c), bboxes (n,4) xyxy in
    network-input pixels, shape = (ih, iw) -> (n,mh,mw) [or (n,ih,iw) when upsample] float32 {0,1}."""
    return process_mask_batch(protos, masks_in, bboxes, None, shape, upsample)


def process_mask_native(protos, masks_in, bboxes, shape):
    """Reference signature (utils/segment/general.py:55-76): up-sample the un-padded prototype window to `shape`, then crop."""
    return process_mask_batch(protos, masks_in, bboxes, None, shape, native=True)
