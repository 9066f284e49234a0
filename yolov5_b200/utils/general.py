"""Post-processing entry points with the reference's signatures (reference utils/general.py:574-581,613-626,658-767).

``non_max_suppression`` runs the whole batch in liby5b200 (y5_nms_batched): one launch set, one device->host read
(the per-image counts) instead of the reference's per-image Python loop with ~20 tiny kernels and several implicit
synchronisations per image.  Indices are bit-exact w.r.t. the reference semantics (tests/test_nms_*.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from .._lib import NmsParams

MAX_WH = 7680.0   # reference utils/general.py:690
MAX_NMS = 30000   # reference utils/general.py:691

_ws_cache: dict = {}


def _iou_threshold_f32(thr: float) -> float:
    """torchvision's CPU nms kernel compares the fp32 IoU against the threshold as a double.  For an fp32 ratio r,
    r > thr(double)  <=>  r > (largest fp32 value <= thr); return that fp32 value."""
    f = np.float32(thr)
    if float(f) > float(thr):
        f = np.nextafter(f, np.float32(-np.inf))
    return float(f)


def _workspace(nbytes: int, device) -> torch.Tensor:
    # scratch is cached per (device, stream): calls on different streams (e.g. validation on a side stream) never share it
    key = (device.index if device.index is not None else torch.cuda.current_device(), _lib.stream_ptr(device))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def nms_device(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, max_det=300, nm=0):
    """Batched device NMS.  Returns (rows (B,max_det,6+nm) fp32, idx (B,max_det) int64, count (B,) int32), all on
    the device, no synchronisation.  idx = candidate id row*nc + cls of each kept detection."""
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if not (isinstance(prediction, torch.Tensor) and prediction.is_cuda):
        raise RuntimeError("y5b200: non_max_suppression runs on CUDA tensors only (no CPU / PyTorch fallback)")
    if prediction.dim() != 3:
        raise ValueError(f"prediction must be (batch, rows, 5+nc+nm), got {tuple(prediction.shape)}")
    pred = prediction.contiguous()
    bs, n, no = pred.shape
    nc = no - nm - 5
    lib = _lib.lib()
    p = NmsParams()
    p.batch, p.n_rows, p.no, p.nc, p.nm = bs, n, no, nc, nm
    p.dtype = _lib.dtype_code(pred.dtype)
    p.conf_thres, p.iou_thres = float(conf_thres), _iou_threshold_f32(iou_thres)
    p.multi_label, p.agnostic = int(bool(multi_label)), int(bool(agnostic))
    p.max_det, p.max_nms, p.max_wh = int(max_det), MAX_NMS, MAX_WH
    cls_t = None
    if classes is not None:
        cls_t = torch.as_tensor(list(classes), dtype=torch.int32, device=pred.device)
        p.classes, p.n_classes = cls_t.data_ptr(), cls_t.numel()
    else:
        p.classes, p.n_classes = None, 0
    need = lib.y5_nms_workspace_bytes(C.byref(p))
    if need < 0:
        _lib.check(-1, "nms_workspace_bytes")
    ws = _workspace(int(need), pred.device)
    ws_ptr = (ws.data_ptr() + 255) & ~255
    rows = torch.empty(bs, max_det, 6 + nm, dtype=torch.float32, device=pred.device)
    idx = torch.empty(bs, max_det, dtype=torch.int64, device=pred.device)
    count = torch.empty(bs, dtype=torch.int32, device=pred.device)
    with _lib.on(pred.device):
        _lib.check(lib.y5_nms_batched(C.byref(p), pred.data_ptr(), rows.data_ptr(), idx.data_ptr(), count.data_ptr(), ws_ptr,
                                      int(need), C.c_void_p(_lib.stream_ptr(pred.device))), "nms_batched")
    return rows, idx, count


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300, nm=0, return_indices=False):
    """Same contract as reference utils/general.py:658-767: list of (n_i, 6+nm) fp32 tensors [xyxy, conf, cls, masks]."""
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]  # model in validation mode: (inference_out, loss_out)
    if labels and any(len(l) for l in labels):
        return _nms_with_apriori_labels(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det, nm, return_indices)
    rows, idx, count = nms_device(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, nm)
    counts = count.tolist()  # the one device->host synchronisation of the call
    out = [rows[b, :c] for b, c in enumerate(counts)]
    if return_indices:
        return out, [idx[b, :c] for b, c in enumerate(counts)]
    return out


def _nms_with_apriori_labels(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det, nm, return_indices):
    """Autolabelling path (reference utils/general.py:706-712, `val.py --save-hybrid`): image xi's label rows [cls, x, y, w, h]
    are appended to its candidates as boxes with obj = cls-score = 1.  The reference's `torch.cat((x, v), 0)` with an fp32
    `v` promotes that image's candidate rows to fp32 BEFORE the obj*cls product, while the first objectness test (:686) was
    still taken in the input dtype: both are reproduced by running such an image on its own in fp32 with the rows that fail
    the input-dtype test zeroed out.  Rare path: one NMS call per labelled image."""
    bs, n, no = prediction.shape
    nc = no - nm - 5
    outs, idxs = [], []
    for xi in range(bs):
        p = prediction[xi : xi + 1]
        lb = labels[xi] if xi < len(labels) else None
        if lb is not None and len(lb):
            lb = torch.as_tensor(lb, device=p.device).float()
            keep = p[..., 4] > conf_thres                      # :686, evaluated in the input dtype
            p32 = p.float()
            p32[..., 4] = torch.where(keep, p32[..., 4], torch.zeros((), device=p.device))
            v = torch.zeros(1, len(lb), no, device=p.device)
            v[0, :, :4] = lb[:, 1:5]
            v[0, :, 4] = 1.0
            v[0, torch.arange(len(lb), device=p.device), lb[:, 0].long() + 5] = 1.0
            p = torch.cat((p32, v), 1)
        rows, idx, count = nms_device(p, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, nm)
        c = int(count[0])
        outs.append(rows[0, :c])
        idxs.append(idx[0, :c])
    return (outs, idxs) if return_indices else outs


def xywh2xyxy(x):
    """[cx, cy, w, h] -> [x1, y1, x2, y2] in the input dtype (ultralytics.utils.ops.xywh2xyxy, used at
    reference utils/general.py:722, val.py:304).  Pure indexing helper for callers; the NMS kernel has its own copy."""
    y = torch.empty_like(x) if isinstance(x, torch.Tensor) else np.empty_like(x)
    half = x[..., 2:4] / 2
    y[..., 0:2] = x[..., 0:2] - half
    y[..., 2:4] = x[..., 0:2] + half
    return y


def xyxy2xywh(x):
    """reference utils/general.py:574-581."""
    y = x.clone() if isinstance(x, torch.Tensor) else np.copy(x)
    y[..., 0] = (x[..., 0] + x[..., 2]) / 2
    y[..., 1] = (x[..., 1] + x[..., 3]) / 2
    y[..., 2] = x[..., 2] - x[..., 0]
    y[..., 3] = x[..., 3] - x[..., 1]
    return y


def clip_boxes(boxes, shape):
    if isinstance(boxes, torch.Tensor):
        boxes[..., 0].clamp_(0, shape[1]); boxes[..., 1].clamp_(0, shape[0])
        boxes[..., 2].clamp_(0, shape[1]); boxes[..., 3].clamp_(0, shape[0])
    else:
        boxes[..., [0, 2]] = boxes[..., [0, 2]].clip(0, shape[1])
        boxes[..., [1, 3]] = boxes[..., [1, 3]].clip(0, shape[0])
    return boxes


def scale_meta(img1_shape, img0_shapes, ratio_pads=None) -> torch.Tensor:
    """(B,5) fp32 [gain, pad_x, pad_y, w0, h0] per image: the numbers reference scale_boxes (utils/general.py:613-621) derives
    from the letterboxed shape img1 and each original shape (or takes from the dataloader's ratio_pad)."""
    rows = []
    for i, s0 in enumerate(img0_shapes):
        rp = ratio_pads[i] if ratio_pads is not None else None
        if rp is None:
            gain = min(img1_shape[0] / s0[0], img1_shape[1] / s0[1])
            pad = (img1_shape[1] - s0[1] * gain) / 2, (img1_shape[0] - s0[0] * gain) / 2
        else:
            gain, pad = rp[0][0], rp[1]
        rows.append([float(gain), float(pad[0]), float(pad[1]), float(s0[1]), float(s0[0])])
    return torch.tensor(rows, dtype=torch.float32).view(-1, 5)


def scale_boxes_batch(rows: torch.Tensor, count, meta: torch.Tensor) -> torch.Tensor:
    """In place, whole batch, one launch: rows (B, max_det, >=4) fp32 with xyxy first (the layout nms_device returns), count
    (B,) int32 valid rows per image (None = all), meta (B,5) from scale_meta."""
    if not rows.is_cuda:
        raise RuntimeError("y5b200: scale_boxes runs on CUDA tensors only (no CPU / PyTorch fallback)")
    assert rows.dtype == torch.float32 and rows.dim() == 3 and rows.stride(2) == 1 and rows.stride(0) == rows.shape[1] * rows.stride(1)
    dev = rows.device
    meta = meta.to(dev, torch.float32).contiguous()
    cnt = count.to(dev, torch.int32).contiguous() if count is not None else None
    with _lib.on(dev):
        _lib.check(_lib.lib().y5_scale_boxes(rows.data_ptr(), rows.stride(1), rows.shape[0] * rows.shape[1], None, rows.shape[1],
                                             cnt.data_ptr() if cnt is not None else None, meta.data_ptr(), C.c_void_p(_lib.stream_ptr(dev))),
                   "scale_boxes")
    return rows


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None):
    """Rescale xyxy boxes from the letterboxed img1_shape back to img0_shape and clip, IN PLACE, like the reference
    (utils/general.py:613-626); `boxes` is an (n, >=4) fp32 CUDA tensor or view (e.g. `det[:, :4]`)."""
    if not (isinstance(boxes, torch.Tensor) and boxes.is_cuda):
        raise RuntimeError("y5b200: scale_boxes runs on CUDA tensors only (no CPU / PyTorch fallback)")
    if boxes.numel() == 0:
        return boxes
    if boxes.dtype != torch.float32 or boxes.dim() != 2 or boxes.stride(1) != 1 or boxes.shape[1] < 4:
        raise TypeError("y5b200: scale_boxes expects an (n, >=4) float32 tensor / view with unit column stride")
    meta = scale_meta(img1_shape, [img0_shape], [ratio_pad]).to(boxes.device)
    with _lib.on(boxes.device):
        _lib.check(_lib.lib().y5_scale_boxes(boxes.data_ptr(), boxes.stride(0), boxes.shape[0], None, 0, None, meta.data_ptr(),
                                             C.c_void_p(_lib.stream_ptr(boxes.device))), "scale_boxes")
    return boxes
