"""Metric matching of the validation loop on the device (reference utils/metrics.py:224-265 process_batch, box_iou as used
at :158,:252, and the per-image loop of val.py:282-318): IoU, class test and the detection<->label matching rule for the
WHOLE batch in one launch, with no `.cpu()` round trip per image."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib


def box_iou(box1: torch.Tensor, box2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """Pairwise IoU of xyxy boxes: (N,4) x (M,4) -> (N,M) fp32, computed by y5_box_iou on the device."""
    if not (box1.is_cuda and box2.is_cuda):
        raise RuntimeError("y5b200: box_iou runs on CUDA tensors only (no CPU / PyTorch fallback)")
    a = box1.float().contiguous()
    b = box2.float().contiguous()
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=a.device)
    if out.numel():
        with _lib.on(a.device):
            _lib.check(_lib.lib().y5_box_iou(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], float(eps), out.data_ptr(),
                                             C.c_void_p(_lib.stream_ptr(a.device))), "box_iou")
    return out


def match_batch(det_rows: torch.Tensor, count, labels6: torch.Tensor, iouv: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """det_rows (B, max_det, >=6) fp32 [x1,y1,x2,y2,conf,cls,...] in native pixels (what scale_boxes produced), count (B,)
    int32 valid rows per image or None; labels6 (nt,6) [img, cls, x1,y1,x2,y2] native pixels -> correct (B, max_det, niou)
    bool on the device."""
    if not det_rows.is_cuda:
        raise RuntimeError("y5b200: process_batch runs on CUDA tensors only (no CPU / PyTorch fallback)")
    dev = det_rows.device
    assert det_rows.dtype == torch.float32 and det_rows.dim() == 3 and det_rows.stride(2) == 1 and det_rows.shape[2] >= 6
    b, max_det = det_rows.shape[:2]
    iouv = iouv.to(dev, torch.float32).contiguous()
    labels6 = labels6.to(dev, torch.float32).contiguous().view(-1, 6)
    correct = torch.empty(b, max_det, iouv.numel(), dtype=torch.uint8, device=dev)
    if b == 0 or max_det == 0:
        return correct.bool()
    cnt = count.to(dev, torch.int32).contiguous() if count is not None else None
    with _lib.on(dev):
        _lib.check(_lib.lib().y5_match_batch(det_rows.data_ptr(), det_rows.stride(0), det_rows.stride(1), cnt.data_ptr() if cnt is not None else None,
                                             b, max_det, labels6.data_ptr() if labels6.numel() else None, labels6.shape[0], iouv.data_ptr(),
                                             iouv.numel(), float(eps), correct.data_ptr(), C.c_void_p(_lib.stream_ptr(dev))), "match_batch")
    return correct.view(torch.bool)


def process_batch(detections, labels, iouv, pred_masks=None, gt_masks=None, overlap=False, masks=False):
    """Reference signature (utils/metrics.py:224): detections (N,6) [x1,y1,x2,y2,conf,cls], labels (M,5) [cls,x1,y1,x2,y2],
    iouv thresholds -> correct (N, len(iouv)) bool on iouv.device.  Box matching only (masks=True is the mask-IoU branch of
    segment/val.py, outside this path)."""
    if masks:
        raise NotImplementedError("y5b200: mask-IoU matching (segment/val.py) is outside the engine's hot path")
    n = detections.shape[0]
    dev = detections.device
    lab6 = torch.cat((torch.zeros(labels.shape[0], 1, device=labels.device, dtype=labels.dtype), labels), 1)
    if n == 0:
        return torch.zeros(0, iouv.numel(), dtype=torch.bool, device=iouv.device)
    det = detections.float().contiguous()[None]
    return match_batch(det, None, lab6, iouv)[0].to(iouv.device)


def labels_to_native(targets: torch.Tensor, meta: torch.Tensor) -> torch.Tensor:
    """val.py:303-306 for all images at once: targets (nt,6) [img, cls, cx, cy, w, h] in network-input pixels, meta (B,5)
    [gain, pad_x, pad_y, w0, h0] -> (nt,6) [img, cls, x1, y1, x2, y2] in native pixels."""
    if not targets.is_cuda:
        raise RuntimeError("y5b200: labels_to_native runs on CUDA tensors only (no CPU / PyTorch fallback)")
    t = targets.float().contiguous().view(-1, 6)
    out = torch.empty_like(t)
    if t.shape[0]:
        m = meta.to(t.device, torch.float32).contiguous()
        with _lib.on(t.device):
            _lib.check(_lib.lib().y5_labels_native(t.data_ptr(), t.shape[0], m.data_ptr(), out.data_ptr(), C.c_void_p(_lib.stream_ptr(t.device))),
                       "labels_native")
    return out


def val_batch_metrics(rows: torch.Tensor, count: torch.Tensor, targets: torch.Tensor, im_shape, shapes, iouv: torch.Tensor):
    """The metric part of val.py:282-318 for a whole batch, on the device: rows/count from nms_device (rows (B,max_det,6+nm)
    in network-input pixels), targets (nt,6) [img, cls, cx, cy, w, h] already in network-input pixels (val.py:274),
    im_shape = (height, width) of the network input, shapes[i] = ((h0, w0), ((ratio_h, ratio_w), (pad_w, pad_h))) as the
    reference dataloader yields.  Returns (predn rows in native space, correct (B,max_det,niou) bool); nothing is synced."""
    from .general import scale_meta, scale_boxes_batch

    meta = scale_meta(im_shape, [s[0] for s in shapes], [s[1] if len(s) > 1 else None for s in shapes]).to(rows.device)
    predn = rows.clone()
    scale_boxes_batch(predn, count, meta)
    labelsn = labels_to_native(targets, meta)
    return predn, match_batch(predn, count, labelsn, iouv)
